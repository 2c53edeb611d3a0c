#!/bin/bash
# ncu --set full capture of the 4-D sweep kernel on named probe cases (one launch each) + a timing pass.
# Usage (GPU box, repo root): bash scripts/gpu_ncu_mech2.sh <tag> <mode> case1 [case2...]
TAG=$1; MODE=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for c in "$@"; do
  PYRODP_MECH2=$MODE timeout 600 python scripts/probe_perf.py $c 2>> $OUT/probe.err | tee -a $OUT/probe_$MODE.jsonl
done
for c in "$@"; do
  PYRODP_MECH2=$MODE timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_mech2 -s 2 -c 1 -f -o $OUT/prof_${c}_$MODE \
      python scripts/probe_perf.py $c > $OUT/ncu_${c}_$MODE.log 2>&1; echo "ncu $c rc=$?"
done
ls -la $OUT
