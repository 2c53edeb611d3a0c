#!/bin/bash
# A/B alternative builds of libpyrodp.so (pyro_b200/_variants/*.so) on the probe cases.
# Usage: bash scripts/gpu_variants.sh <tag> <cases...>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== default"; timeout 600 python scripts/probe_perf.py "$@" 2>&1 | tee $OUT/probe_default.jsonl
for so in pyro_b200/_variants/*.so; do
  echo "== $so"; PYRODP_LIB=$PWD/$so timeout 600 python scripts/probe_perf.py "$@" 2>&1 | tee $OUT/probe_$(basename $so .so).jsonl
done
