#!/bin/bash
# The very last GPU visit of round 2 (about two GPU-minutes left): the spline / rollout tests once more (the spline class
# now takes its tables from the device builder), then the timing probe of the two additions.
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
timeout 40 python -m pytest tests/test_zz_after_the_sweep_gpu.py -m gpu -x -q > $OUT/pytest_zz.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" | tee -a $OUT/pytest_zz.log; tail -3 $OUT/pytest_zz.log
timeout 60 python scripts/probe_after_sweep.py > $OUT/probe_after_sweep.jsonl 2> $OUT/probe_after_sweep.err; echo "probe rc=$? t=$(( $(date +%s) - T0 ))s"; cat $OUT/probe_after_sweep.jsonl; tail -3 $OUT/probe_after_sweep.err
