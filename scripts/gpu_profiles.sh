#!/bin/bash
# Final ncu captures of the shipped sweep kernels at BASELINE sizes + launch list of the bench command.
# Usage (GPU box, repo root): bash scripts/gpu_profiles.sh <tag>
TAG=${1:-r02g}
OUT=gpurun_out/$TAG; mkdir -p $OUT
EXTRA=sm__inst_executed_pipe_fp64.sum,l1tex__data_pipe_lsu_wavefronts.sum
for c in cfg2 cfg3 cfg4 dp81; do
  K=regex:sweep_
  timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k $K -s 3 -c 1 -f -o $OUT/prof_$c \
      python scripts/probe_perf.py $c > $OUT/ncu_$c.log 2>&1; echo "ncu $c rc=$?"
done
LIGHT=dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 1200 ncu --metrics $LIGHT --clock-control none -k regex:sweep_mech2 -s 1 -c 1 -f -o $OUT/prof_cfg5_light \
    python scripts/probe_perf.py cfg5 > $OUT/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
echo "== launch list of the bench command (cfg4 as the workload keeps it short; same code path)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_cfg4.csv \
    python bench.py --steps 5 --warmup 3 --workload cfg4 --sub-workloads cfg2,cfg3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
ls -la $OUT
