#!/bin/bash
# A/B of the halo exchange (peer-memory stores vs NCCL send/recv) under torchrun.  Usage: bash scripts/gpu_halo_ab.sh <tag> <N>
TAG=$1; N=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
for H in ${HALOS:-peer nccl peer nccl}; do
  PYRODP_HALO=$H timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2> $OUT/err_$H.log | grep '"metric"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$H', d['config']['parallelism'], 'value %.4e' % d['value'], 'ms/step', round(d['ms_per_step'], 4), 'by rank', [round(x, 4) for x in d['ms_per_step_by_rank']], 'wall', round(d['wall_s_timed_region'] * 1e3 / d['steps'], 4))
" | tee -a $OUT/halo_ab.txt
done
