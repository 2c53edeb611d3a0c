#!/bin/bash
# ncu of the spline table sweep (one launch, full set) and the launch list of one spline backup (fit kernels vs sweep).
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
timeout 40 ncu --set full --clock-control none --import-source on -k regex:sweep_lut_spline -s 2 -c 1 -f -o $OUT/prof_spline1001 \
    python scripts/probe_after_sweep.py spline1001 > $OUT/ncu_spline.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 25 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spline -c 24 --csv --log-file $OUT/launches_spline1001.csv \
    python scripts/probe_after_sweep.py spline1001 > $OUT/launches.log 2>&1; echo "launch list rc=$? t=$(( $(date +%s) - T0 ))s"; tail -4 $OUT/launches_spline1001.csv
