#!/bin/bash
# Spline evaluation with tabulated reciprocals: parity tests of the file + one timing line (the round's last GPU seconds).
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 9 python scripts/probe_after_sweep.py spline1001 > $OUT/probe_spline.jsonl 2> $OUT/probe_spline.err; echo "probe rc=$?"; cat $OUT/probe_spline.jsonl
timeout 14 python -m pytest tests/test_zz_after_the_sweep_gpu.py -m gpu -x -q -k spline > $OUT/pytest_zz.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_zz.log; tail -3 $OUT/pytest_zz.log
