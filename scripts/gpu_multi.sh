#!/bin/bash
# Multi-GPU visit: parity under torchrun + bench at N ranks.  Usage: bash scripts/gpu_multi.sh <tag> <N> [workload]
TAG=$1; N=$2; WL=${3:-cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
echo "== multigpu_check"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > $OUT/multigpu_check.log 2>&1; echo "check rc=$?"; grep "\[multigpu\]" $OUT/multigpu_check.log | tail -30; grep -i "error\|Traceback" -A5 $OUT/multigpu_check.log | head -30
echo "== bench N=$N $WL"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --workload $WL > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; grep '"metric"' $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
echo "== bench N=1 $WL"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --workload $WL --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; grep '"metric"' $OUT/bench_n1.json
