#!/bin/bash
# N-GPU visit without the single-GPU tail: parity under torchrun + bench at N ranks.  Usage: bash scripts/gpu_multi8.sh <tag> <N>
TAG=$1; N=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
echo "== multigpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > $OUT/multigpu_check.log 2>&1; echo "check rc=$?"; grep -c "OK$" $OUT/multigpu_check.log; grep -E "ALL OK|MISMATCH|FAIL|Traceback" $OUT/multigpu_check.log | head
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench rc=$?"; grep '"metric"' $OUT/bench_n$N.json | cut -c1-400; tail -3 $OUT/bench_n$N.err
echo "== reference arm under torchrun"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | grep '"impl"' | cut -c1-200
