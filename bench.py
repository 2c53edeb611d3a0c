#!/usr/bin/env python
"""bench.py — Bellman-sweep throughput of the B200 grid-DP engine (state-action evals/s).

    python bench.py --gpus 1 --steps 20 --warmup 3            # own arm (CUDA kernels)
    python bench.py --impl reference --steps 3 --warmup 1     # reference arm (CPU, oracle port)
    torchrun ... bench.py --gpus N ...                        # N>1: one rank per GPU, NCCL

A "step" is ONE Bellman sweep (one pass of the hot path over the whole grid): for every node
and every action, Euler step, n-linear interpolation of J, stage cost, min/argmin, plus the
fused convergence reduction (and, for N>1, the all-gather of the new J over NVLink).
Metric, config and roofline definitions: BASELINE.json / SURVEY.md section 8(d) / DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PI = float(np.pi)
# BASELINE.json configs (SURVEY.md 8d "Concrete synthetic inputs")
WORKLOADS = {
    "cfg1": dict(system="SinglePendulum", x_grid_dim=[51, 51], u_grid_dim=[11], xbar=[-3.14, 0.0], INF=300.0, dt=0.05),
    "cfg2": dict(system="SinglePendulum", x_grid_dim=[1001, 1001], u_grid_dim=[201], xbar=[-3.14, 0.0], INF=300.0, dt=0.05),
    "cfg3": dict(system="TwoLinkManipulator", x_grid_dim=[101] * 4, u_grid_dim=[21, 21], INF=1000.0, dt=0.05),
    "cfg4": dict(system="CartPole", x_grid_dim=[151] * 4, u_grid_dim=[51], xbar=[0.0, PI, 0.0, 0.0], INF=1000.0, dt=0.05),
    "cfg5": dict(system="DoublePendulum", x_grid_dim=[201] * 4, u_grid_dim=[31, 31], dt=0.1,
                 x_lb=[-5.0, -1.5, -4.0, -4.0], x_ub=[0.5, 4.0, 5.5, 7.0], u_lb=[-12.0, -12.0], u_ub=[12.0, 12.0],
                 xbar=[0.0, 0.0, 0.0, 0.0], Q=[1.0, 0.5, 0.1, 0.05], R=[0.05, 0.05], INF=1000.0, EPS=1.0),
}
L2_FLUSH_BYTES = 256 << 20  # > 126 MB L2


def b_eval(n, A):
    """Algorithmic bytes per eval, the contract figure of SURVEY.md 8(d): 8*2^n + 24/A."""
    return 8.0 * (1 << n) + 24.0 / A


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the sweep kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


def issue_model(system_id, evals, kernel_ms, clocks):
    """What actually bounds the sweep: an FP64 warp instruction holds a sub-partition's issue
    port for two cycles (scripts/micro/fp64_peak.cu, profiles/r01_fp64_peak_micro.txt), so a warp costs
    2*FP64 + other instructions.  Instruction counts per warp-eval are ncu's (source page of profiles/r01K);
    the measured cycles come from this run's kernel time."""
    if system_id != 1:
        return {"resource": "issue port + L1 data pipe (see DESIGN.md section 5, profiles/r01b_cp101.txt)"}
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    measured = kernel_ms * 1e-3 * mhz * 1e6 * sms * 4 / (evals / 32.0)
    f64, other = 19.45, 18.63
    return {"resource": "issue port: FP64 warp instructions take 2 issue cycles", "fp64_inst_per_warp_eval": f64,
            "other_inst_per_warp_eval": other, "model_cycles_per_warp_eval": 2 * f64 + other,
            "measured_cycles_per_warp_eval": measured, "frac_of_issue_limit": (2 * f64 + other) / measured,
            "arithmetic_floor_cycles_per_warp_eval": 2 * 17.5 + 3.0,
            "source": "profiles/r01K_cfg2 (ncu source page counts of the final kernel), profiles/r01_fp64_peak_micro.txt"}


def weak_scaled(case, world):
    """Per-GPU work fixed: axis 0 carries world x the planes of the named configuration."""
    case = dict(case)
    dims = list(case["x_grid_dim"])
    dims[0] = dims[0] * world
    case["x_grid_dim"] = dims
    return case


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: NVML polled every ~2 ms (nvidia-smi's 100 ms loop is
    too coarse for a timed region of a few milliseconds) — in a helper PROCESS, so that the benchmark process
    itself never opens an NVML session or runs a polling thread beside its own CUDA calls."""

    def __init__(self, gpu_index):
        self.proc, self.ok = None, False
        try:
            self.proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--clock-sampler", str(gpu_index)],
                                         stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.ok = self.proc.stdout.readline().strip() == "ready"
        except Exception:
            self.ok = False

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.ok:
            return out
        try:
            self.proc.stdin.close()             # end of the timed region
            line = self.proc.stdout.readline()
            self.proc.wait(timeout=5)
            out.update(json.loads(line))
        except Exception:
            pass
        return out


def clock_sampler_main(gpu_index):
    """Helper process: poll NVML until stdin closes, then print one JSON summary."""
    import select
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
    except Exception:
        print("unavailable", flush=True)
        return
    names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
             nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
    samples, reasons = [], set()
    print("ready", flush=True)
    while True:
        try:
            samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            for bit, name in names.items():
                if mask & bit:
                    reasons.add(name)
        except Exception:
            pass
        if select.select([sys.stdin], [], [], 0.002)[0]:   # EOF on stdin = stop
            break
    sm = [x[0] for x in samples] or [0.0]
    print(json.dumps({"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": max_mhz,
                      "power_w_max": float(max(x[1] for x in samples)) if samples else None,
                      "reasons": sorted(reasons), "samples": len(samples)}), flush=True)


# --------------------------------------------------------------------------------------------------
# CPU baselines (oracle port) — rank 0 only
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: the reference's LUT sweep (dynamicprogramming.py:564-570, scipy RGI) on a node range."""
    case, lo, hi, reps = args
    from oracle import np_oracle as npo
    from tests.cases import oracle_objects
    grid, cost = oracle_objects(case)
    J_next = np.random.default_rng(0).uniform(0, 250, grid.N)
    x_next, _, _, G = grid.tables(cost, lo, hi)  # one-off table build, NOT timed (reference builds them once)
    t0 = time.perf_counter()
    for _ in range(reps):
        npo.lut_sweep(grid.x_level, grid.dims, J_next, x_next, G, 1.0, use_scipy=True)
    return time.perf_counter() - t0


def cpu_reference_rate(case, n_procs, nodes_per_proc, reps):
    """evals/s of the NumPy/SciPy LUT sweep port with n_procs processes on a bounded node sample."""
    import multiprocessing as mp
    from tests.cases import oracle_objects
    grid, _ = oracle_objects(case)
    nodes_per_proc = min(nodes_per_proc, grid.N // n_procs)
    jobs = [(case, i * nodes_per_proc, (i + 1) * nodes_per_proc, reps) for i in range(n_procs)]
    t0 = time.perf_counter()
    if n_procs == 1:
        times = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            times = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    evals = n_procs * nodes_per_proc * grid.A * reps
    return evals / max(times), evals, wall


def cpu_native_rate(case, n_nodes, reps):
    """evals/s of the C/OpenMP restatement (oracle/dp_oracle.c), all host threads, on-the-fly dynamics."""
    from oracle import c_oracle
    from pyro_b200 import problem
    from tests.cases import build_case
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    n_nodes = min(n_nodes, P.N)
    J_next = np.random.default_rng(0).uniform(0, 250, P.N)
    c_oracle.sweep_fused(P, J_next, 0, min(4096, n_nodes))
    t0 = time.perf_counter()
    for _ in range(reps):
        c_oracle.sweep_fused(P, J_next, 0, n_nodes)
    dt = time.perf_counter() - t0
    return n_nodes * P.A * reps / dt, c_oracle.n_threads_default()


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference_arm(args, case, wl_name):
    """--impl reference: the reference's CPU path (oracle port: NumPy + SciPy RGI LUT sweep) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    from tests.cases import oracle_objects
    grid, _ = oracle_objects(case)
    n, A = grid.spec.n, grid.A
    # bounded sample: ~2e6 evals per process per step keeps one step at a fraction of a second
    nodes_per_proc = max(1, int(2.0e6 // A))
    for _ in range(max(args.warmup, 0)):
        cpu_reference_rate(case, cores, nodes_per_proc, 1)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, evals, _ = cpu_reference_rate(case, cores, nodes_per_proc, 1)
        rates.append(r)
    wall = time.perf_counter() - t0
    value = float(np.median(rates))
    native, native_threads = cpu_native_rate(case, 1 << 15, 2)
    sample = (f"{cores} processes x {nodes_per_proc} nodes x {A} actions per step of the same grid "
              f"(tables prebuilt, LUT sweep only, scipy RGI; dynamicprogramming.py:564-570)")
    line = {
        "impl": "reference", "metric": "state_action_evals_per_s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, **{k: case[k] for k in ("system", "x_grid_dim", "u_grid_dim")},
                   "note": "CPU reference path; each step = a bounded node sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "cpu_baseline_native": {"value": native, "unit": "evals/s", "cores": native_threads, "kind": "port",
                                "sample": "C/OpenMP restatement (oracle/dp_oracle.c), 32768 nodes x all actions, on-the-fly dynamics"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "roofline": None,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--clock-sampler":
        return clock_sampler_main(int(sys.argv[2]))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base = WORKLOADS[args.workload]
    case = weak_scaled(base, world) if (args.scaling == "weak" and world > 1) else dict(base)
    wl_name = {"cfg1": "SinglePendulum 51x51 x 11 actions", "cfg2": "SinglePendulum 1001x1001 state grid x 201 actions",
               "cfg3": "TwoLinkManipulator 101^4 x 21^2 actions", "cfg4": "CartPole 151^4 x 51 actions",
               "cfg5": "DoublePendulum 201^4 x 31^2 actions"}[args.workload] + f" (BASELINE {args.workload})"
    if world > 1 and args.scaling == "weak":
        wl_name += f", axis 0 x{world} (per-GPU slab = the named grid)"

    if args.impl == "reference":
        run_reference_arm(args, base, wl_name)
        return

    import torch
    import torch.distributed as dist
    from pyro_b200 import problem
    from pyro_b200.distributed import ShardedEngine
    from pyro_b200.engine import Engine
    from tests.cases import build_case

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    _, grid, cf = build_case(case)
    n, A, N = grid.sys.n, grid.actions_n, grid.nodes_n
    torch.cuda.set_stream(torch.cuda.Stream())   # a real stream, not the legacy default one: torch events and the engine share it
    stream = torch.cuda.current_stream()
    if world > 1:
        eng = ShardedEngine(grid, cf, 1.0)
        kernel_eng = eng.eng
    else:
        eng = Engine(problem.extract(grid, cf, 1.0))
        eng.set_stream(stream.cuda_stream)
        kernel_eng = eng
    eng.eval_terminal_cost()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.sweep(1)
    barrier()
    evals_per_step = float(N) * A

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    e2e = None
    # (measured before the clock sampler starts, so that no NVML polling runs beside the host<->device pipeline.
    #  On these virtualised hosts the same call varies 0.52-0.68 ms from process to process, the first process on a
    #  fresh box being the slow one: profiles/r01w_e2e_sampler.txt, profiles/r01o_e2e_chunks.jsonl)
    if not args.no_e2e:
        # every rank holds the full host J (the reference API's array) but uploads only the planes it
        # keeps (slab + halo) and reads back only its own slab of J and pi
        slab_n = kernel_eng.slab_nodes
        J_host = torch.empty(N, dtype=torch.float64).pin_memory()
        Js_host = torch.empty(slab_n, dtype=torch.float64).pin_memory()
        pis_host = torch.empty(slab_n, dtype=torch.int64).pin_memory()
        J_host.zero_()
        J_np, Js_np, pis_np = J_host.numpy(), Js_host.numpy(), pis_host.numpy()
        kernel_eng.get_J(Js_np)
        lo = kernel_eng.slab_begin * kernel_eng.plane
        J_np[lo:lo + slab_n] = Js_np
        if world > 1:  # complete the host copy once (not timed) so every rank uploads real halo data
            t = torch.from_numpy(J_np).cuda()
            dist.all_reduce(t)
            J_host.copy_(t.cpu())
        held = (kernel_eng.alloc_end - kernel_eng.alloc_begin) * kernel_eng.plane
        n_e2e = max(5, min(args.steps, 20))

        # ONE C-ABI call per step with host arrays on both sides (pdp_sweep_host): H2D of the J_next planes the rank
        # holds (slab + halo of the full host array), the sweep, D2H of the slab's J and pi — pipelined over plane
        # chunks and replayed as one CUDA graph.  A rank needs no halo exchange for a single host-to-host sweep.
        def e2e_step():
            kernel_eng.sweep_host(J_np, Js_np, pis_np)
        for _ in range(3):
            e2e_step()
        barrier()
        e2e_steps_ms = []
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            t1 = time.perf_counter()
            e2e_step()               # blocking call: returns when J and pi are in the host buffers
            e2e_steps_ms.append(1e3 * (time.perf_counter() - t1))
        barrier()
        dt = time.perf_counter() - t0
        h2d, d2h = 8.0 * held, 16.0 * slab_n
        if world > 1:
            t = torch.tensor([dt, -dt, h2d, d2h], device="cuda", dtype=torch.float64)
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt, h2d, d2h = float(mx[0].item()), float(t[2].item()), float(t[3].item())
        if world > 1:
            eng.set_J(J_np)   # device-resident sweeps continue from a J whose halo planes are current
        e2e = {"value": evals_per_step * n_e2e / dt, "unit": "evals/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e,
               "step_ms_min_median_max": [float(np.min(e2e_steps_ms)), float(np.median(e2e_steps_ms)), float(np.max(e2e_steps_ms))],
               "call": "pdp_sweep_host (H2D J_next -> sweep -> D2H J, pi; chunk-pipelined, one CUDA graph), pinned host buffers"
                       + (" (per rank: the planes it holds up, its slab down)" if world > 1 else "")}

    # ---- timed region: K sweeps, L2 flushed before each, device time by CUDA events ----------------
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("BENCH_NO_SAMPLER")) else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = kernel_eng.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    for s, e in ev:
        flush.zero_()            # write 256 MB > L2: the next sweep re-reads J_next from HBM
        s.record()
        eng.sweep_nowait()       # one Bellman sweep incl. the fused dJ statistics (+ halo exchange for N>1), enqueued
        e.record()               # asynchronously: the host never waits inside the timed region
    last_stats = eng.collect_stats()  # the K statistics triples (one small D2H; all-reduced over ranks for N>1)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = np.array([s.elapsed_time(e) for s, e in ev])
    total_ms = float(step_ms.sum())
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        allt = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allt, torch.tensor([total_ms], device="cuda", dtype=torch.float64))
        per_rank_ms = [float(x.item()) / args.steps for x in allt]
        total_ms = max(per_rank_ms) * args.steps
    launches = kernel_eng.launch_count - launches0
    clocks = sampler.stop() if sampler else None

    value = evals_per_step * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel alone: back-to-back launches on the stream, events around the batch --------
    kb = max(args.steps, 5)
    if world == 1:
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        k0.record()
        eng.sweep(kb)
        k1.record()
        torch.cuda.synchronize()
        kernel_ms = k0.elapsed_time(k1) / kb
    else:
        kernel_ms = total_ms / args.steps
    peak, peak_src = measured_peak()
    slab_evals = evals_per_step / world
    achieved = slab_evals * b_eval(n, A) / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload), "peak_source": peak_src,
                "kernel": {1: "sweep_pendulum_kernel", 2: "sweep_mech2_kernel<TWOLINK>", 3: "sweep_mech2_kernel<CARTPOLE>"}[kernel_eng.problem.system_id],
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_eval": b_eval(n, A),
                "algorithmic_bytes_per_launch": slab_evals * b_eval(n, A),
                "compulsory_dram_bytes_per_launch": 24.0 * N / world,
                "dram": {"achieved": (ncu_traffic(args.workload) or 24.0 * N / world) / (kernel_ms * 1e-3) / 1e9, "unit": "GB/s",
                         "frac": (ncu_traffic(args.workload) or 24.0 * N / world) / (kernel_ms * 1e-3) / 1e9 / peak,
                         "note": "measured DRAM bytes per launch (ncu) over this run's kernel time: the sweep is not HBM-bound"},
                "binding_resource": issue_model(kernel_eng.problem.system_id, slab_evals, kernel_ms, clocks),
                "note": "contract figure of SURVEY 8(d): the 2^n-corner J gather is served by L1/L2, so DRAM traffic is ~24 B/node "
                        "and frac can exceed 1; the binding resource is the FP64 pipe (see DESIGN.md, profiles/)"}

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------------------
    cpu = cpu_nat = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nodes = max(1, int(4.0e6 // A))
        rate, evals, wall = cpu_reference_rate(base, 1, nodes, 5)
        cpu = {"value": rate, "unit": "evals/s", "cores": 1, "kind": "port",
               "sample": f"reference LUT sweep (NumPy + scipy RGI, dynamicprogramming.py:564-570) on {nodes} nodes x {A} actions "
                         f"of the same grid, 5 passes, tables prebuilt; the reference is single-threaded"}
        nat, thr = cpu_native_rate(base, 1 << 16, 3)
        cpu_nat = {"value": nat, "unit": "evals/s", "cores": thr, "kind": "port",
                   "sample": "C/OpenMP restatement oracle/dp_oracle.c, 65536 nodes x all actions x 3 passes, on-the-fly dynamics"}

    if rank == 0:
        line = {
            "metric": "state_action_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "system": case["system"], "x_grid_dim": case["x_grid_dim"],
                       "u_grid_dim": case["u_grid_dim"], "dt": case["dt"], "alpha": 1.0, "nodes": N, "actions": A,
                       "evals_per_step": evals_per_step,
                       "parallelism": (f"slab{world}/{eng.mode}/{eng.halo}" + ("+overlap" if eng.overlap else "")) if world > 1 else "single",
                       "l2": f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)", "J0": "h(x) then warm-up sweeps"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "cpu_baseline_native": cpu_nat,
            "wall_s_timed_region": t_wall, "ms_per_step_by_rank": per_rank_ms, "step_ms_min_max": [float(step_ms.min()), float(step_ms.max())],
            "last_sweep_stats": {"j_max": float(last_stats[-1][0]), "delta_max": float(last_stats[-1][1]), "delta_min": float(last_stats[-1][2])},
            "J_Linf_error": "see tests/test_parity_gpu.py (bit-exact vs reference goldens)",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
