#!/bin/bash
# Multi-GPU visit: GPU test-suite (incl. torchrun parity of every exchange mode) + bench at N ranks.
# Usage: bash scripts/gpu_multi.sh <tag> <N> [workload] [steps]
TAG=$1; N=$2; WL=${3:-cfg5}; STEPS=${4:-3}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== multigpu_check"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > $OUT/multigpu_check.log 2>&1; echo "check rc=$?"; grep "\[multigpu\]" $OUT/multigpu_check.log | tail -60; grep -i "error\|Traceback" -A8 $OUT/multigpu_check.log | head -40
echo "== pytest multi engine"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "multi_engine or planner_uses" > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_multi.log
echo "== bench N=$N $WL"; (time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 3 --workload $WL > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err); echo "bench rc=$?"; grep '"metric"' $OUT/bench_n$N.json | cut -c1-1200; tail -8 $OUT/bench_n$N.err
