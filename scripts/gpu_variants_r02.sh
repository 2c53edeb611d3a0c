#!/bin/bash
# A/B of register budgets of the 4-D range kernel (blocks per SM forced through __launch_bounds__) + micro-benchmarks.
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "" mb5 mb6; do
  lib=""; [ -n "$v" ] && lib=$PWD/pyro_b200/_variants/libpyrodp_$v.so
  echo "== variant ${v:-shipped (4 blocks/SM)}"
  PYRODP_LIB=$lib timeout 600 python scripts/probe_perf.py cfg3 cfg4 dp81 2>> $OUT/err.log | sed "s/^{/{\"variant\": \"${v:-mb4}\", /" | tee -a $OUT/variants.jsonl
done
echo "== micro: fp64_peak"; ./scripts/micro/fp64_peak | tee $OUT/fp64_peak_micro.txt
echo "== micro: l1_gather"; ./scripts/micro/l1_gather | tee $OUT/l1_gather_micro.txt
