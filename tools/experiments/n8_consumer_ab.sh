# 8 GPUs, small steps: where the consumer sits in the step graph (own branch / behind the record packing)
for cfg in 2 3; do for c in branch side; do timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $cfg --consumer $c --no-subconfigs --no-cpu-baseline --steps 300 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg$cfg $c', round(r['ms_per_step']*1e3,2), 'fused', round(r['fused_step']['ms_per_step']*1e3,2), 'w/o', round(r['ms_per_step_without_exchange']*1e3,2))"; done; done
