for cfg in 5 2 4; do python bench.py --config $cfg --no-subconfigs --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=r['kernels']
print('C$cfg', '| step', round(r['ms_per_step']*1e3,1), 'fused', round(r['fused_step']['ms_per_step']*1e3,1), '| enc', round(k['encode']['ms']*1e3,1), 'dec', round(k['decode_expected']['ms']*1e3,1), 'loss', round(k['loss_fwd_bwd']['ms']*1e3,1), 'dark', round(k['decode_dark']['ms']*1e3,1), r['parity_check'].get('ok'))"; done
