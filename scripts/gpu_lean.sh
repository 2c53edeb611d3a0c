#!/bin/bash
# Lean gpurun visit: GPU parity tests, smoke, bench (+reference arm), probe, ncu launch list of the bench command.
# Usage: bash scripts/gpu_lean.sh <tag> [probe cases...]
TAG=${1:-lean}; shift
CASES=${@:-cfg1 cfg2 tl61 cp101 dp61}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -4 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
echo "== probe"; timeout 900 python scripts/probe_perf.py $CASES > $OUT/probe.jsonl 2> $OUT/probe.err; cat $OUT/probe.jsonl; tail -3 $OUT/probe.err
echo "== ncu launch list (bench command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
