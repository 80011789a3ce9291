# Step time when the decode / loss kernels are capped to fewer CTAs per SM so that the kernels of the step's two streams
# can be resident together (by default each of them fills the register file or the shared memory on its own).
run() { cfg=$1; shift; env "$@" python bench.py --config $cfg --no-subconfigs --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=r['kernels']
print('C$cfg', '$*', '| step', round(r['ms_per_step']*1e3,1), 'fused', round(r['fused_step']['ms_per_step']*1e3,1), '| enc', round(k['encode']['ms']*1e3,1), 'dec', round(k['decode_expected']['ms']*1e3,1), 'loss', round(k['loss_fwd_bwd']['ms']*1e3,1), 'loss_enc', round(k['loss_encoded_fwd_bwd']['ms']*1e3,1), r['parity_check'].get('ok'))"; }
for cfg in 5 2; do
run $cfg X=1
run $cfg PP_DECODE_CTAS=3
run $cfg PP_DECODE_CTAS=2
run $cfg PP_DECODE_CTAS=3 PP_LOSS_CTAS=3
run $cfg PP_DECODE_CTAS=2 PP_LOSS_CTAS=3
run $cfg PP_DECODE_CTAS=1 PP_LOSS_CTAS=3
done
