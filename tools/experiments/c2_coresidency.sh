run() { env "$@" python bench.py --config 2 --no-subconfigs --no-cpu-baseline --steps 300 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=r['kernels']
print('$*', '| step', round(r['ms_per_step']*1e3,1), 'fused', round(r['fused_step']['ms_per_step']*1e3,1), '| enc', round(k['encode']['ms']*1e3,1), 'dec', round(k['decode_expected']['ms']*1e3,1), 'loss', round(k['loss_fwd_bwd']['ms']*1e3,1), 'loss_enc', round(k['loss_encoded_fwd_bwd']['ms']*1e3,1), r['parity_check'].get('ok') if isinstance(r.get('parity_check'),dict) else r.get('parity_check'))"; }
run X=1
run PP_DECODE_GRID=296
run PP_DECODE_GRID=296 PP_LOSS_CTAS=2
run PP_DECODE_GRID=444 PP_LOSS_CTAS=1
run PP_LOSS_CTAS=2
run PP_DECODE_GRID=296 PP_LOSS_CTAS=3
run PP_DECODE_GRID=148 PP_LOSS_CTAS=3
