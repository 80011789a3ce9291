for cfg in 2 5; do for f in "" "--plain-backward"; do python bench.py --config $cfg --no-subconfigs --no-cpu-baseline --steps 200 $f 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C$cfg $f', '| step', round(r['ms_per_step']*1e3,2), 'fused', round(r['fused_step']['ms_per_step']*1e3,2), r['parity_check'].get('ok'), r['parity_check'].get('grad_max_rel'))"; done; done
