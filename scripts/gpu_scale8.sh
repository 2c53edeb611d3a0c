#!/bin/bash
# 8-GPU visit: torchrun parity (all exchange modes, full-size cfg4 hashes), cfg5 strong scaling at N=8 and N=4,
# the single-process multi-GPU engine on cfg4.  Usage: bash scripts/gpu_scale8.sh <tag> [steps]
TAG=${1:-r02f}; STEPS=${2:-3}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt; nvidia-smi topo -m > $OUT/topo.txt 2>&1
for N in 8 4; do
  echo "== bench N=$N cfg5"; (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps $STEPS --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err); echo "bench rc=$?"; grep '"metric"' $OUT/bench_n$N.json | cut -c1-700; tail -4 $OUT/bench_n$N.err
done
echo "== multigpu_check N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_check.py > $OUT/multigpu_check.log 2>&1; echo "check rc=$?"; grep "\[multigpu\]" $OUT/multigpu_check.log | grep -v ": OK" | tail -20; grep -c ": OK" $OUT/multigpu_check.log
echo "== single-process MultiEngine cfg4 on 8 GPUs"; timeout 600 python scripts/probe_multi.py cfg4 8 2>&1 | tail -3 | tee $OUT/probe_multi_cfg4.jsonl
timeout 600 python scripts/probe_multi.py cfg3 8 2>&1 | tail -1 | tee -a $OUT/probe_multi_cfg4.jsonl
echo "== bench N=8 cfg4"; (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg4 > $OUT/bench_cfg4_n8.json 2> $OUT/bench_cfg4_n8.err); echo "bench rc=$?"; grep '"metric"' $OUT/bench_cfg4_n8.json | cut -c1-700
