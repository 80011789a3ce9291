# finalize_kernel as a programmatic dependent launch (PP_LOSS_PDL=1, default) against an ordinary launch
for cfg in 2 5; do for pdl in 1 0 1 0; do PP_LOSS_PDL=$pdl python bench.py --config $cfg --no-subconfigs --no-cpu-baseline --steps 200 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C$cfg PDL=$pdl', '| step', round(r['ms_per_step']*1e3,2), 'fused', round(r['fused_step']['ms_per_step']*1e3,2), r['parity_check'].get('ok'), r['parity_check'].get('loss_value_rel'))"; done; done
