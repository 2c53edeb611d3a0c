#!/usr/bin/env python
"""Per-launch counters of the sweep kernels from ncu reports -> profiles/traffic.json (read by bench.py's roofline block).

    python scripts/ncu_counters.py cfg3=gpurun_out/r02g/prof_cfg3.ncu-rep cfg4=... [--note "..."]

For every workload: dram bytes (read + write), FP64 warp instructions, L1 data-pipe (LSU) wavefronts, warp instructions,
the utilisation percentages ncu reports for the FP64 pipe, the L1 data pipe and the issue slots, the kernel name and time.
Runs here (no GPU needed): it only reads the reports."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
    "fp64_warp_inst": "sm__inst_executed_pipe_fp64.sum", "l1_wavefronts": "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1_wavefronts_shared": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "warp_inst": "smsp__inst_executed.sum",
    "pct_fp64_pipe": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "pct_l1_data_pipe": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "pct_issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "pct_dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "time": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[-1]       # the last captured launch
    out = {"kernel": row[hdr.index("Kernel Name")].split("(")[0]}
    for key, metric in WANT.items():
        if metric in hdr:
            i = hdr.index(metric)
            v = float(row[i].replace(",", ""))
            out[key] = v * UNIT_SCALE.get(units[i], 1.0)
    return out


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table = {k: v for k, v in table.items() if isinstance(v, dict)}
    note = None
    for arg in sys.argv[1:]:
        if arg.startswith("--note="):
            note = arg[7:]
            continue
        wl, rep = arg.split("=", 1)
        c = read(rep)
        rec = {"kernel": c["kernel"], "kernel_ms_under_ncu": 1e3 * c.get("time", 0.0),
               "dram_bytes": c.get("dram_read", 0.0) + c.get("dram_write", 0.0), "dram_read": c.get("dram_read"), "dram_write": c.get("dram_write"),
               "fp64_warp_inst": c.get("fp64_warp_inst"), "l1_wavefronts": c.get("l1_wavefronts"), "l1_wavefronts_shared": c.get("l1_wavefronts_shared"),
               "warp_inst": c.get("warp_inst"),
               "ncu_pct": {"fp64_pipe": c.get("pct_fp64_pipe"), "l1_data_pipe": c.get("pct_l1_data_pipe"), "issue": c.get("pct_issue"), "dram": c.get("pct_dram"),
                           "l1_hit": c.get("l1_hit_pct"), "l2_hit": c.get("l2_hit_pct"), "warps_active": c.get("warps_active_pct")},
               "registers": c.get("regs"), "source": os.path.relpath(os.path.abspath(rep), ROOT) + " (ncu --clock-control none, one launch)"}
        if note:
            rec["note"] = note
        table[wl] = rec
        print(wl, json.dumps(rec)[:400])
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
