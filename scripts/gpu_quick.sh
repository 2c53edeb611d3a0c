#!/bin/bash
# Quick GPU iteration: parity tests + per-config probe (+ optional ncu of named probe cases).
# Usage: bash scripts/gpu_quick.sh <tag> [ncu-case ...]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
PYRODP_LANES=4 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "goldens or mid_size or edge" > $OUT/pytest_gpu_g4.log 2>&1; echo "pytest G=4 rc=$?"; tail -2 $OUT/pytest_gpu_g4.log
PYRODP_LANES=16 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "goldens or mid_size or edge" > $OUT/pytest_gpu_g16.log 2>&1; echo "pytest G=16 rc=$?"; tail -2 $OUT/pytest_gpu_g16.log
timeout 900 python scripts/probe_perf.py cfg1 cfg2 tl61 cp101 dp61 > $OUT/probe.jsonl 2> $OUT/probe.err; cat $OUT/probe.jsonl; tail -3 $OUT/probe.err
for c in "$@"; do
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 3 -c 1 -f -o $OUT/prof_$c \
      python scripts/probe_perf.py $c > $OUT/ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
