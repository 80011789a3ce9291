# sweep of CTAs-per-SM caps for the decoder (side stream) and the loss kernel (main stream)
run() { echo "D=$1 L=$2 $(PP_DECODE_CTAS=$1 PP_LOSS_CTAS=$2 python bench.py --steps 200 --warmup 5 --no-cpu-baseline $3 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in j['kernels'].items()})")"; }
run 0 0; run 0 0 --serial
for d in 2 3 4; do for l in 2 3; do run $d $l; done; done
