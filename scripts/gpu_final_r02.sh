#!/bin/bash
# Last GPU visit of round 2 (about four GPU-minutes were left): parity of the changed / new kernels first, then the A/B of
# the pendulum loops, then one ncu capture of the shipped pendulum kernel.  Every step has its own timeout.
# Usage: bash scripts/gpu_final_r02.sh <tag>
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
echo "== pytest: pendulum parity + rollout + spline"
timeout 100 python -m pytest tests/test_parity_gpu.py tests/test_zz_after_the_sweep_gpu.py -m gpu -x -q \
    -k "(((pend_ or pendulum) and not dpend and not pe_pend) or config2 or edge_cases or rollout or spline) and not True" \
    > $OUT/pytest_new.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" | tee -a $OUT/pytest_new.log; tail -4 $OUT/pytest_new.log
echo "== probe A/B"
timeout 30 python scripts/probe_pend.py nest pair nest > $OUT/probe_pend.jsonl 2> $OUT/probe_pend.err; echo "probe rc=$? t=$(( $(date +%s) - T0 ))s"; cat $OUT/probe_pend.jsonl; tail -2 $OUT/probe_pend.err
PYRODP_LIB=$PWD/pyro_b200/libpyrodp_b9.so timeout 20 python scripts/probe_pend.py nest >> $OUT/probe_pend.jsonl 2>> $OUT/probe_pend.err; echo "probe b9 rc=$? t=$(( $(date +%s) - T0 ))s"; tail -1 $OUT/probe_pend.jsonl
echo "== ncu"
timeout 50 ncu --set full --clock-control none --import-source on -k regex:sweep_pendulum -s 3 -c 1 -f -o $OUT/prof_cfg2 \
    python scripts/probe_pend.py nest --sweeps 3 > $OUT/ncu_cfg2.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - T0 ))s"
echo "== remaining GPU tests while time is left"
timeout 40 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not True and not full_size and not cfg5 and not cfg4 and not cfg3 and not multi" \
    > $OUT/pytest_rest.log 2>&1; echo "pytest rest rc=$? t=$(( $(date +%s) - T0 ))s" | tee -a $OUT/pytest_rest.log; tail -3 $OUT/pytest_rest.log
