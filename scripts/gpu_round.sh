#!/bin/bash
# One gpurun visit: GPU parity tests, bench line, per-config probe, ncu launch list + full captures.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -4 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json
echo "== probe"; timeout 900 python scripts/probe_perf.py cfg1 cfg2 tl61 cp101 dp61 > $OUT/probe.jsonl 2> $OUT/probe.err; cat $OUT/probe.jsonl
echo "== ncu launch list (bench command)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full: pendulum (cfg2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_pendulum -s 3 -c 2 -f -o $OUT/prof_pendulum \
    python scripts/probe_perf.py cfg2 > $OUT/ncu_pendulum.log 2>&1; echo "ncu pend rc=$?"
echo "== ncu full: 4-D (tl61, cp101)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_mech2 -s 3 -c 1 -f -o $OUT/prof_twolink \
    python scripts/probe_perf.py tl61 > $OUT/ncu_twolink.log 2>&1; echo "ncu twolink rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_mech2 -s 3 -c 1 -f -o $OUT/prof_cartpole \
    python scripts/probe_perf.py cp101 > $OUT/ncu_cartpole.log 2>&1; echo "ncu cartpole rc=$?"
ls -la $OUT
