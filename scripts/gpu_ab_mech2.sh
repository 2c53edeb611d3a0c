#!/bin/bash
# A/B of the 4-D sweep kernel variants on one B200: ms per sweep for each PYRODP_MECH2 mode.
# Usage (on the GPU box, repo root): bash scripts/gpu_ab_mech2.sh <tag> [cases...]
TAG=${1:-r02a}; shift
CASES=${@:-"cfg3 cfg4 dp81"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for mode in ${MODES:-generic range}; do
  echo "== PYRODP_MECH2=$mode"
  PYRODP_MECH2=$mode timeout 600 python scripts/probe_perf.py $CASES 2> $OUT/probe_$mode.err | tee $OUT/probe_$mode.jsonl
done
