#!/bin/bash
# A/B visit: parity (default build, all lane variants), variant builds on probe cases, ncu full of named cases.
# Usage: bash scripts/gpu_ab.sh <tag> "<probe cases>" [ncu-case ...]
TAG=$1; CASES=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for L in 1 4 16; do
PYRODP_LANES=$L timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "goldens or mid_size or edge" > $OUT/pytest_gpu_g$L.log 2>&1; echo "pytest G=$L rc=$?"; tail -1 $OUT/pytest_gpu_g$L.log
done
echo "== default"; timeout 600 python scripts/probe_perf.py $CASES 2>&1 | tee $OUT/probe_default.jsonl
for so in pyro_b200/_variants/*.so; do
  echo "== $so"; PYRODP_LIB=$PWD/$so timeout 600 python scripts/probe_perf.py $CASES 2>&1 | tee $OUT/probe_$(basename $so .so).jsonl
done
for c in "$@"; do
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 3 -c 1 -f -o $OUT/prof_$c \
      python scripts/probe_perf.py $c > $OUT/ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
