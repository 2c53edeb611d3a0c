#!/usr/bin/env python
"""bench.py — Bellman-sweep throughput of the B200 grid-DP engine (state-action evals/s).

    python bench.py --gpus 1 --steps 20 --warmup 5            # own arm (CUDA kernels), BASELINE cfg5, 1 GPU
    torchrun ... bench.py --gpus N ...                        # N>1: one rank per GPU, same grid sharded (strong scaling)
    python bench.py --impl reference --steps 3 --warmup 1     # reference arm: the unmodified pyro classes on the host cores

A "step" is ONE Bellman sweep (one pass of the hot path over the whole grid): for every node and every action, Euler
step, n-linear interpolation of J, stage cost, min/argmin, plus the fused convergence reduction and, for N>1, the halo
exchange of the new J over NVLink.  The workload is BASELINE.json's configs[4] — DoublePendulum 201^4 state grid x 31^2
actions, the configuration the metric's "1/2/4/8 B200" is quoted on (39 GB of state, fits one GPU) — strong-scaled:
the SAME grid on 1, 2, 4 or 8 GPUs, slabs over the outermost axis.  At N=1 the JSON line also carries sub-records for
configs 2, 3 and 4.  Metric, config and roofline definitions: BASELINE.json / SURVEY.md section 8(d) / DESIGN.md.
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
WORKLOAD_NAMES = {"cfg1": "SinglePendulum 51x51 x 11 actions", "cfg2": "SinglePendulum 1001x1001 state grid x 201 actions",
                  "cfg3": "TwoLinkManipulator 101^4 x 21^2 actions", "cfg4": "CartPole 151^4 x 51 actions",
                  "cfg5": "DoublePendulum 201^4 x 31^2 actions"}
# what the unmodified reference can build in about half a minute of its Python table loops (discretizer.py:342-376): the
# same system, bounds, dt, cost and ACTION grid on a coarser state grid (the named grids need 3 GB ... 50 TB of tables)
REFERENCE_SAMPLE_DIMS = {"cfg1": [51, 51], "cfg2": [55, 55], "cfg3": [6, 6, 6, 6], "cfg4": [11, 11, 11, 11], "cfg5": [5, 5, 5, 5]}
L2_FLUSH_BYTES = 256 << 20  # > 126 MB L2
FP64_WARP_INST_PER_CLK_PER_SM = 2.0  # measured: scripts/micro/fp64_peak.cu, profiles/r01_fp64_peak_micro.txt (64 FP64 lanes per SM)


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


def ncu_counters(workload):
    """Per-launch counters of the sweep kernel from the committed ncu captures (profiles/traffic.json), or {}:
    dram_bytes (dram__bytes_read.sum + dram__bytes_write.sum), fp64_warp_inst (sm__inst_executed_pipe_fp64.sum),
    l1_wavefronts (l1tex__data_pipe_lsu_wavefronts.sum), warp_inst (smsp__inst_executed.sum), ncu_pct, source
    (written by scripts/ncu_counters.py)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(p)).get(workload)
        return rec if isinstance(rec, dict) else ({"dram_bytes": rec} if rec else {})
    except Exception:
        return {}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: NVML polled in a helper PROCESS (every 2 ms for short regions,
    50 ms for long ones), so that the benchmark process itself never opens an NVML session beside its own CUDA calls."""

    def __init__(self, gpu_index, period_s=0.002):
        self.proc, self.ok = None, False
        try:
            self.proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--clock-sampler", str(gpu_index), str(period_s)],
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


def clock_sampler_main(gpu_index, period_s):
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
        if select.select([sys.stdin], [], [], period_s)[0]:   # EOF on stdin = stop
            break
    sm = [x[0] for x in samples] or [0.0]
    print(json.dumps({"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": max_mhz,
                      "power_w_max": float(max(x[1] for x in samples)) if samples else None,
                      "reasons": sorted(reasons), "samples": len(samples)}), flush=True)


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# --------------------------------------------------------------------------------------------------
# CPU baselines: the oracle port (own arm's cpu_baseline) and the unmodified reference (reference arm)
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: the reference's LUT sweep (dynamicprogramming.py:564-570, scipy RGI) on a node range."""
    case, lo, hi, reps = args
    from oracle import np_oracle as npo
    from tests.cases import oracle_objects
    grid, cost = oracle_objects(case)
    J_next = np.random.default_rng(0).uniform(0, 250, grid.N)
    x_next, _, _, G = grid.tables(cost, lo, hi)  # one-off table build, NOT timed (the reference builds them once)
    t0 = time.perf_counter()
    for _ in range(reps):
        npo.lut_sweep(grid.x_level, grid.dims, J_next, x_next, G, 1.0, use_scipy=True)
    return time.perf_counter() - t0


def cpu_port_rate(case, n_procs, nodes_per_proc, reps):
    """evals/s of the NumPy/SciPy LUT-sweep port with n_procs processes, each on a bounded node sample of the box."""
    import multiprocessing as mp
    A = int(np.prod(case["u_grid_dim"]))
    nodes = max(1, min(nodes_per_proc, _sample_nodes(case)))
    jobs = [(dict(case, x_grid_dim=_sample_dims(case)), 0, nodes, reps)] * n_procs
    t0 = time.perf_counter()
    if n_procs == 1:
        times = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            times = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    evals = n_procs * nodes * A * reps
    return evals / max(times), evals, wall, nodes


def _sample_dims(case):
    """Grid the port's J_next is drawn on: the named grid when it is small, else the same box at <= 61 levels per axis
    (the port evaluates x_next / RGI per node, so only the level spacing changes, not the work per eval)."""
    return [min(int(d), 61 if len(case["x_grid_dim"]) == 4 else 1001) for d in case["x_grid_dim"]]


def _sample_nodes(case):
    return int(np.prod(np.array(_sample_dims(case), dtype=np.int64)))


def cpu_native_rate(case, n_nodes, reps):
    """evals/s of the C/OpenMP restatement (oracle/dp_oracle.c), all host threads, on-the-fly dynamics."""
    from oracle import c_oracle
    from pyro_b200 import problem
    from tests.cases import build_case
    _, grid, cf = build_case(dict(case, x_grid_dim=_sample_dims(case)))
    P = problem.extract(grid, cf, 1.0)
    n_nodes = min(n_nodes, P.N)
    J_next = np.random.default_rng(0).uniform(0, 250, P.N)
    c_oracle.sweep_fused(P, J_next, 0, min(4096, n_nodes))
    t0 = time.perf_counter()
    for _ in range(reps):
        c_oracle.sweep_fused(P, J_next, 0, n_nodes)
    dt = time.perf_counter() - t0
    return n_nodes * P.A * reps / dt, c_oracle.n_threads_default()


def _reference_worker(args):
    """One process = one replica of the UNMODIFIED reference (pyro from baseline/_ref or /root/reference): build the grid
    and its look-up tables with the reference's own Python loops (not timed), then time whole sweeps —
    initialize_backward_step + compute_backward_step + finalize_backward_step (dynamicprogramming.py:175-261, :557-570)."""
    case, warm, steps, lut, barrier = args
    from oracle import ref_loader
    ns = ref_loader.load()
    t0 = time.perf_counter()
    with ref_loader.quiet():
        _, grid, _, dp = ref_loader.build_reference(ns, case, lut=lut)
    build_s = time.perf_counter() - t0
    dp.save_time_history = False
    times = []
    with ref_loader.quiet():
        for _ in range(warm):
            dp.initialize_backward_step(); dp.compute_backward_step(); dp.finalize_backward_step()
        if barrier is not None:
            barrier.wait()
        for _ in range(steps):
            t1 = time.perf_counter()
            dp.initialize_backward_step(); dp.compute_backward_step(); dp.finalize_backward_step()
            times.append(time.perf_counter() - t1)
    return times, build_s, int(grid.nodes_n), int(grid.actions_n)


def reference_rate(case, n_procs, warm, steps, lut=True):
    """evals/s of n_procs replicas of the unmodified reference, one per host core: (rate, ms per step, build s, N, A)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    if n_procs == 1:
        res = [_reference_worker((case, warm, steps, lut, None))]
    else:
        barrier = ctx.Manager().Barrier(n_procs)
        with ctx.Pool(n_procs) as pool:
            res = pool.map(_reference_worker, [(case, warm, steps, lut, barrier)] * n_procs)
    N, A = res[0][2], res[0][3]
    per_step = np.max(np.array([r[0] for r in res]), axis=0)       # slowest replica of each step
    rate = n_procs * N * A / float(np.median(per_step))
    return rate, 1e3 * float(np.median(per_step)), float(max(r[1] for r in res)), N, A


def run_reference_arm(args, case, wl_key, wl_name):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_loader
    cores = host_cores()
    A = int(np.prod(case["u_grid_dim"]))
    extra = {}
    if ref_loader.available():
        # the real classes: DynamicProgrammingWithLookUpTable over GridDynamicSystem, one replica per host core
        sample_case = dict(case, x_grid_dim=REFERENCE_SAMPLE_DIMS[wl_key])
        rate, ms, build_s, N, A = reference_rate(sample_case, cores, max(args.warmup, 1), args.steps, lut=True)
        kind = "reference"
        sample = (f"{cores} replicas (one per host core; the reference is single-threaded) of the UNMODIFIED "
                  f"pyro.planning.dynamicprogramming.DynamicProgrammingWithLookUpTable (from {os.path.relpath(ref_loader.ref_root(), ROOT) if ref_loader.ref_root().startswith(ROOT) else ref_loader.ref_root()}) "
                  f"on the same system, bounds, dt, cost and {case['u_grid_dim']} action grid with a {REFERENCE_SAMPLE_DIMS[wl_key]} state grid "
                  f"({N} nodes x {A} actions per replica per step; the named grid's tables do not fit or finish); whole sweeps "
                  f"(initialize + compute + finalize_backward_step), tables prebuilt in {build_s:.0f} s (not timed)")
        if not args.no_extra:
            # the two other reference measurements BASELINE.md names: cfg1 at full size, LUT class and base class
            c1 = WORKLOADS["cfg1"]
            r1, ms1, b1, N1, A1 = reference_rate(c1, 1, 1, 3, lut=True)
            extra["cfg1_lut_1core"] = {"value": r1, "unit": "evals/s", "ms_per_sweep": ms1, "table_build_s": b1, "nodes": N1, "actions": A1,
                                       "what": "DynamicProgrammingWithLookUpTable.compute_backward_step, SinglePendulum 51x51x11, 1 process"}
            r0, ms0, _, _, _ = reference_rate(c1, 1, 0, 1, lut=False)
            extra["cfg1_base_class_1core"] = {"value": r0, "unit": "evals/s", "ms_per_sweep": ms0,
                                              "what": "DynamicProgramming.compute_backward_step (per-pair sys.f, dynamicprogramming.py:195-236), 1 process"}
    else:
        nodes_per_proc = max(1, int(2.0e6 // A))
        for _ in range(max(args.warmup, 0)):
            cpu_port_rate(case, cores, nodes_per_proc, 1)
        rates = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rates.append(cpu_port_rate(case, cores, nodes_per_proc, 1)[0])
        ms = 1e3 * (time.perf_counter() - t0) / max(args.steps, 1)
        rate, kind = float(np.median(rates)), "port"
        sample = (f"reference not installed (baseline/_ref absent): oracle port, {cores} processes x {nodes_per_proc} nodes x {A} actions "
                  f"per step, LUT sweep with scipy RGI (dynamicprogramming.py:564-570), tables prebuilt")
    native, native_threads = cpu_native_rate(case, 1 << 15, 2)
    line = {
        "impl": "reference", "metric": "state_action_evals_per_s", "value": rate, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, **{k: case[k] for k in ("system", "x_grid_dim", "u_grid_dim")},
                   "note": "CPU reference path; each step = one whole sweep of a bounded sample of the workload on every host core"},
        "cpu_baseline": {"value": rate, "unit": "evals/s", "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_native": {"value": native, "unit": "evals/s", "cores": native_threads, "kind": "port",
                                "sample": "C/OpenMP restatement (oracle/dp_oracle.c), 32768 nodes x all actions, on-the-fly dynamics"},
        "reference_extra": extra,
        "e2e": {"value": rate, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "roofline": None,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """torch / distributed handles shared by the workload runs of one process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_stream(torch.cuda.Stream())   # a real stream, not the legacy default one: torch events and the engine share it
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        self.sms = torch.cuda.get_device_properties(self.local_rank).multi_processor_count

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        """all-reduce a list of floats over the ranks (max / sum); identity on one GPU."""
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM}[op])
        return [float(x) for x in t.tolist()]

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)


def pinned(ctx, n, dtype):
    """Pinned host array of n elements (falls back to pageable memory if the box refuses to pin that much)."""
    t = ctx.torch.empty(int(n), dtype=dtype)
    try:
        return t.pin_memory().numpy(), True
    except Exception:
        return t.numpy(), False


def parity_sample(ctx, keng, P, n_random=24, width=256):
    """Sampled check of the LAST sweep against the C oracle (oracle/dp_oracle.c — the checker, never the thing measured):
    the engine still holds the sweep's input (J_next, halo planes included) and its outputs, so no extra sweep is needed.
    Ranges: random ones in this rank's slab, its first and last nodes (the slab seams for N>1) and the slab's middle.
    Returns max |J - J_ref| / max |J_ref| (north_star tolerance 1e-5), the pi mismatch count and the nodes checked."""
    from oracle import c_oracle
    plane = keng.plane
    lo_node, hi_node = keng.slab_begin * plane, keng.slab_end * plane
    held_lo, held = keng.alloc_begin * plane, (keng.alloc_end - keng.alloc_begin) * plane
    J_next = np.empty(P.N)                       # virtual: only the planes this rank holds are ever touched
    keng.get_range("J_next", held_lo, held, out=J_next[held_lo:held_lo + held])
    rng = np.random.default_rng(1234 + ctx.rank)
    exhaustive = float(hi_node - lo_node) * P.A <= 5e8          # cfg2-sized slabs: the oracle does every node in a second or two
    width = (hi_node - lo_node) if exhaustive else min(width, hi_node - lo_node)
    starts = [int(s) for s in rng.integers(lo_node, hi_node - width + 1, n_random)]
    starts += [lo_node, hi_node - width, (lo_node + hi_node - width) // 2]
    if exhaustive:
        starts = [lo_node]
    err_abs, ref_max, mism, checked, exact = 0.0, 0.0, 0, 0, True
    for s in starts:
        Jr, pr = c_oracle.sweep_fused(P, J_next, s, s + width)
        Jg, pg = keng.get_range("J", s, width), keng.get_range("pi", s, width)
        fin = np.isfinite(Jr) & np.isfinite(Jg)
        err_abs = max(err_abs, float(np.abs(Jg[fin] - Jr[fin]).max()) if fin.any() else 0.0)
        mism += int((pg != pr).sum()) + int((np.isfinite(Jr) != np.isfinite(Jg)).sum())
        ref_max = max(ref_max, float(np.abs(Jr[fin]).max()) if fin.any() else 0.0)
        exact = exact and bool(np.array_equal(Jg, Jr))
        checked += width
    err_abs, ref_max = ctx.reduce([err_abs], "max")[0], ctx.reduce([ref_max], "max")[0]
    mism, checked, inexact = (int(v) for v in ctx.reduce([mism, checked, 0 if exact else 1], "sum"))
    return {"J_Linf_error": err_abs / ref_max if ref_max > 0 else err_abs, "J_Linf_abs": err_abs, "pi_mismatches": mism,
            "nodes_checked": checked, "ranges": len(starts) * ctx.world, "bit_exact": inexact == 0,
            "against": "oracle/dp_oracle.c backup of the same J_next, " + ("EVERY node of the slab" if exhaustive else "sampled node ranges incl. slab seams")}


def run_workload(ctx, args, wl_key, primary):
    """Time one BASELINE configuration on the GPUs of this job; returns the record (rank 0) of that workload."""
    torch = ctx.torch
    from pyro_b200 import problem
    from pyro_b200.distributed import ShardedEngine
    from pyro_b200.engine import Engine
    from tests.cases import build_case
    world, rank = ctx.world, ctx.rank
    case = dict(WORKLOADS[wl_key])
    _, grid, cf = build_case(case)
    n, A, N = grid.sys.n, grid.actions_n, grid.nodes_n
    evals_per_step = float(N) * A
    stream = torch.cuda.current_stream()
    t_setup = time.perf_counter()
    if world > 1:
        # slabs whose boundaries equalise the measured compute time per rank (one calibration sweep, not timed)
        eng = ShardedEngine(grid, cf, 1.0) if os.environ.get("BENCH_EQUAL_SLABS") else ShardedEngine.balanced(grid, cf, 1.0)
        keng = eng.eng
    else:
        eng = Engine(problem.extract(grid, cf, 1.0))
        eng.set_stream(stream.cuda_stream)
        keng = eng
    P = keng.problem
    eng.eval_terminal_cost()
    steps = args.steps if primary else max(3, min(args.steps, args.sub_steps))
    warmup = args.warmup if primary else 3

    # ---- warm-up, and the per-step estimate that sizes the optional parts ------------------------------------
    w0, w1 = ctx.event(), ctx.event()
    ctx.barrier()
    w0.record()
    eng.sweep(warmup)
    w1.record()
    ctx.barrier()
    est_ms = ctx.reduce([w0.elapsed_time(w1) / max(warmup, 1)], "max")[0]
    long_steps = est_ms > 1000.0

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # ONE C-ABI call per step with host arrays on both sides (pdp_sweep_host / pdp_sweep_host_local): H2D of the
        # J_next planes the rank holds (slab + halo), the sweep, D2H of the slab's J and pi — pipelined over plane chunks
        # (one CUDA graph when the buffers are pinned).  A rank needs no halo exchange for a single host-to-host sweep.
        plane = keng.plane
        held = (keng.alloc_end - keng.alloc_begin) * plane
        slab_n = keng.slab_nodes
        (J_held, p1), (Js, p2), (pis, p3) = pinned(ctx, held, torch.float64), pinned(ctx, slab_n, torch.float64), pinned(ctx, slab_n, torch.int64)
        keng.get_range("J", keng.alloc_begin * plane, held, out=J_held)
        n_warm, n_e2e = (1, 3) if long_steps else (3, max(5, min(steps, 20)))

        def e2e_step():
            keng.sweep_host_local(J_held, Js, pis)     # blocking: returns when J and pi are in the host buffers
        for _ in range(n_warm):
            e2e_step()
        ctx.barrier()
        step_ms = []
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            t1 = time.perf_counter()
            e2e_step()
            step_ms.append(1e3 * (time.perf_counter() - t1))
        ctx.barrier()
        dt = ctx.reduce([time.perf_counter() - t0], "max")[0]
        h2d, d2h = ctx.reduce([8.0 * held, 16.0 * slab_n], "sum")
        e2e = {"value": evals_per_step * n_e2e / dt, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": n_e2e, "warmup": n_warm, "ms_per_step": 1e3 * dt / n_e2e, "pinned": bool(p1 and p2 and p3),
               "step_ms_min_median_max": [float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms))],
               "call": "pdp_sweep_host_local (H2D J_next planes -> sweep -> D2H J, pi; chunk-pipelined, one CUDA graph), pinned host "
                       "buffers" + (" (per rank: the planes it holds up, its slab down)" if world > 1 else "")}
        if world > 1:
            eng.eng.exchange_current()   # device-resident sweeps continue from a J whose halo planes are current
        del J_held, Js, pis

    # ---- timed region: K sweeps, L2 flushed before each, device time by CUDA events ----------------
    sampler = ClockSampler(ctx.local_rank, 0.05 if long_steps else 0.002) if (rank == 0 and not os.environ.get("BENCH_NO_SAMPLER")) else None
    ev = [(ctx.event(), ctx.event()) for _ in range(steps)]
    launches0 = keng.launch_count
    ctx.barrier()
    t_wall0 = time.perf_counter()
    for s, e in ev:
        ctx.flush.zero_()        # write 256 MB > L2: the next sweep re-reads J_next from HBM
        s.record()
        eng.sweep_nowait()       # one Bellman sweep incl. the fused dJ statistics (+ halo exchange for N>1), enqueued
        e.record()               # asynchronously: the host never waits inside the timed region
    last_stats = eng.collect_stats()  # the K statistics triples (one small D2H; all-reduced over ranks for N>1)
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = np.array([s.elapsed_time(e) for s, e in ev])
    total_ms = float(step_ms.sum())
    per_rank_ms = [total_ms / steps]
    if world > 1:
        allt = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        ctx.dist.all_gather(allt, torch.tensor([total_ms], device="cuda", dtype=torch.float64))
        per_rank_ms = [float(x.item()) / steps for x in allt]
        total_ms = max(per_rank_ms) * steps
    launches = int(ctx.reduce([keng.launch_count - launches0], "sum")[0])
    clocks = sampler.stop() if sampler else None
    value = evals_per_step * steps / (total_ms * 1e-3)

    # ---- sampled parity of the last timed sweep (numbers, not a pointer to the tests) ------------------------
    parity = parity_sample(ctx, keng, P) if not args.no_parity else None

    # ---- the dominant kernel alone; for N>1 also the exchange alone ------------------------------------------
    exchange_ms = None
    if world == 1:
        if long_steps:
            kernel_ms = float(np.median(step_ms))          # one launch per step: the step IS the kernel
        else:
            kb = max(steps, 5)
            torch.cuda.synchronize()
            k0, k1 = ctx.event(), ctx.event()
            ctx.flush.zero_()
            k0.record()
            eng.sweep(kb)
            k1.record()
            torch.cuda.synchronize()
            kernel_ms = k0.elapsed_time(k1) / kb
    else:
        km, xm = [], []
        for _ in range(2 if long_steps else 5):
            a, b, c, d = ctx.event(), ctx.event(), ctx.event(), ctx.event()
            ctx.barrier()
            a.record(); keng.sweep_async(); b.record(); keng.commit_sweep()     # the slab's planes in one launch, no exchange
            torch.cuda.synchronize()
            ctx.barrier()
            c.record(); keng.exchange_current(); d.record()                     # the halo planes alone
            torch.cuda.synchronize()
            km.append(a.elapsed_time(b)); xm.append(c.elapsed_time(d))
        kernel_ms = ctx.reduce([float(np.median(km))], "max")[0]
        exchange_ms = ctx.reduce([float(np.median(xm))], "max")[0]

    # ---- roofline: the unit that binds (ncu, profiles/) — the FP64 pipe for the 2-D kernel, the L1 data pipe for the 4-D
    #      gathers — evaluated on THIS run's kernel time; the SURVEY 8(d) HBM contract figure and DRAM traffic beside it ----
    peak_hbm, peak_src = measured_peak()
    slab_evals = evals_per_step / world
    ncu = ncu_counters(wl_key)
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    clk = ctx.sms * mhz * 1e6                                         # SM-cycles per second
    kernel_s = kernel_ms * 1e-3

    def unit(count, per_clk_per_sm, label, source):
        if not count:
            return None
        ach, peak = (count / world) / kernel_s / 1e9, per_clk_per_sm * clk / 1e9
        return {"achieved": ach, "peak": peak, "unit": label, "frac": ach / peak, "per_launch": count / world, "peak_source": source}
    fp64 = unit(ncu.get("fp64_warp_inst"), FP64_WARP_INST_PER_CLK_PER_SM, "G FP64 warp-inst/s",
                f"{FP64_WARP_INST_PER_CLK_PER_SM} FP64 warp-inst/clk/SM (scripts/micro/fp64_peak.cu, profiles/r01_fp64_peak_micro.txt) x {ctx.sms} SMs x {mhz:.0f} MHz")
    l1 = unit(ncu.get("l1_wavefronts"), 1.0, "G L1 data-pipe wavefronts/s",
              f"1 LSU wavefront/clk/SM (ncu l1tex__data_pipe_lsu_wavefronts peak) x {ctx.sms} SMs x {mhz:.0f} MHz")
    cands = [(k, v) for k, v in (("fp64_issue", fp64), ("l1_data_pipe", l1)) if v]
    bound, top = max(cands, key=lambda kv: kv[1]["frac"]) if cands else ("fp64_issue", None)
    contract = slab_evals * b_eval(n, A) / kernel_s / 1e9
    dram = ncu.get("dram_bytes")
    roofline = {
        "bound": bound, "achieved": top["achieved"] if top else None, "peak": top["peak"] if top else None,
        "unit": top["unit"] if top else None, "frac": top["frac"] if top else None,
        "traffic": dram / world if dram else None,
        "kernel": keng.kernel_info, "kernel_ms": kernel_ms, "exchange_ms": exchange_ms,
        "fp64_issue": fp64, "l1_data_pipe": l1,
        "fp64_inst_per_eval": ncu["fp64_warp_inst"] * 32.0 / evals_per_step if ncu.get("fp64_warp_inst") else None,
        "l1_wavefronts_per_warp_eval": ncu["l1_wavefronts"] * 32.0 / evals_per_step if ncu.get("l1_wavefronts") else None,
        "ncu_pct": ncu.get("ncu_pct"), "counters_source": ncu.get("source"), "counters_kernel": ncu.get("kernel"),
        "traffic_over_compulsory": (dram / (24.0 * N)) if dram else None,
        "compulsory_dram_bytes_per_launch": 24.0 * N / world,
        "dram": {"achieved": (dram / world) / kernel_s / 1e9, "peak": peak_hbm, "unit": "GB/s",
                 "frac": (dram / world) / kernel_s / 1e9 / peak_hbm} if dram else None,
        "contract": {"bound": "hbm", "achieved": contract, "peak": peak_hbm, "unit": "GB/s", "frac": contract / peak_hbm,
                     "peak_source": peak_src, "algorithmic_bytes_per_eval": b_eval(n, A),
                     "algorithmic_bytes_per_launch": slab_evals * b_eval(n, A),
                     "note": "SURVEY 8(d) contract figure: the 2^n-corner J gather counted as memory traffic; it is served by L1/L2, "
                             "so this fraction can exceed 1 and is not a bandwidth statement"},
        "note": "counters per launch come from the committed ncu capture of the same kernel and workload (profiles/traffic.json); "
                "time, clock and therefore every rate and fraction are this run's",
        "bound_note": ("fp64_issue = FP64 warp instructions per second over the measured 2.0 / clk / SM; the FP64 unit is a pipe shared by the "
                       "SM's four sub-partitions, and the r02p A/B (11 % fewer non-FP64 instructions, same sweep time) shows that pipe, not the "
                       "issue port, binds the 2-D kernel") if bound == "fp64_issue" else None,
    }

    # ---- CPU baseline beside it (rank 0, N=1, primary workload only) -----------------------------------------
    cpu = cpu_nat = None
    if primary and rank == 0 and world == 1 and not args.no_cpu_baseline:
        nodes = max(1, int(4.0e6 // A))
        rate, evals, wall, used = cpu_port_rate(case, 1, nodes, 5)
        cpu = {"value": rate, "unit": "evals/s", "cores": 1, "kind": "port",
               "sample": f"reference LUT sweep port (NumPy + scipy RGI, dynamicprogramming.py:564-570) on {used} nodes x {A} actions "
                         f"of the same box on a {_sample_dims(case)} grid, 5 passes, tables prebuilt; the reference is single-threaded "
                         f"(the unmodified reference itself is timed by --impl reference)"}
        nat, thr = cpu_native_rate(case, 1 << 16, 3)
        cpu_nat = {"value": nat, "unit": "evals/s", "cores": thr, "kind": "port",
                   "sample": "C/OpenMP restatement oracle/dp_oracle.c, 65536 nodes x all actions x 3 passes, on-the-fly dynamics"}

    parallelism = "single"
    if world > 1:
        parallelism = f"slab{world} over axis 0/{eng.mode}/{eng.halo}" + ("+overlap" if eng.overlap else "") + \
                      f", halo {keng.halo_lo}+{keng.halo_hi} planes of {keng.plane * 8 / 1e6:.1f} MB, slab boundaries {eng.bounds}" + \
                      (" (rebalanced from measured per-rank compute times)" if getattr(eng, "calibration", {}).get("rebalanced") else " (equal)")
    rec = {
        "metric": "state_action_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD_NAMES[wl_key]} (BASELINE {wl_key})" + (f", the same grid sharded over {world} GPUs" if world > 1 else ""),
                   "system": case["system"], "x_grid_dim": case["x_grid_dim"], "u_grid_dim": case["u_grid_dim"], "dt": case["dt"],
                   "alpha": 1.0, "nodes": N, "actions": A, "evals_per_step": evals_per_step, "parallelism": parallelism,
                   "l2": f"flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)", "J0": "h(x) then warm-up sweeps"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "cpu_baseline": cpu, "cpu_baseline_native": cpu_nat,
        "J_Linf_error": parity["J_Linf_error"] if parity else None, "pi_mismatches": parity["pi_mismatches"] if parity else None,
        "parity": parity,
        "wall_s_timed_region": t_wall, "ms_per_step_by_rank": per_rank_ms, "step_ms_min_max": [float(step_ms.min()), float(step_ms.max())],
        "last_sweep_stats": {"j_max": float(last_stats[-1][0]), "delta_max": float(last_stats[-1][1]), "delta_min": float(last_stats[-1][2])},
        "setup_s": time.perf_counter() - t_setup,
    }
    eng.close()
    torch.cuda.empty_cache()
    return rec


def after_the_sweep_records(n_traj=131072, policy_dims=(201, 201), policy_sweeps=100, spline_case=None):
    """N = 1 only, after the timed workloads: the two rows that sit next to the sweep — closed-loop
    Euler rollout batches (pdp_rollout) and the bicubic-spline table sweep (pdp_set_interpolant) — each timed through the
    public API and checked against its fixture of the unmodified reference (tests/golden/).  Never raises."""
    out = {}
    try:
        from pyro_b200 import dynamicprogramming
        from tests.cases import CASES, build_case
        golden = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        # ---- rollouts: parity on the fixture's grid, throughput on a 201 x 201 x 21 policy --------------------------------
        gold = golden("rollout_pend_51x51x11")
        _, grid, cf = build_case(CASES["pend_51x51x11"])
        dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
        dp.verbose = False
        dp.compute_steps(int(gold["sweeps"]))
        npts, tf = int(gold["npts"]), float(gold["tf"])
        _, x, u = dp.compute_closed_loop_trajectories(gold["x0"], tf, npts)
        rec = {"parity": {"against": "tests/golden/rollout_pend_51x51x11.npz: (ctl + sys).compute_trajectory(tf, n, 'euler') of the unmodified reference",
                          "policy_equal": bool(np.array_equal(dp.pi, gold["pi"])),
                          "x_Linf_error": float(np.abs(x - gold["x"]).max()), "u_Linf_error": float(np.abs(u - gold["u"]).max())}}
        sys_, grid, cf = build_case(dict(CASES["pend_51x51x11"], x_grid_dim=list(policy_dims), u_grid_dim=[21]))
        dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
        dp.verbose = False
        dp.compute_steps(policy_sweeps)
        B, npts = int(n_traj), 1001
        x0 = np.random.default_rng(1).uniform(np.asarray(sys_.x_lb) * 0.9, np.asarray(sys_.x_ub) * 0.9, (B, 2))
        dp.compute_closed_loop_trajectories(x0[:min(B, 1024)], 10.0, npts, stride=100)
        t0 = time.perf_counter()
        dp.compute_closed_loop_trajectories(x0, 10.0, npts, stride=100)
        wall = time.perf_counter() - t0
        rec.update({"workload": f"SinglePendulum {policy_dims[0]} x {policy_dims[1]} x 21 policy after {policy_sweeps} sweeps, {B} trajectories x {npts} points "
                                "(tf = 10), every 100th point copied back",
                    "wall_s": wall, "value": B * (npts - 1) / wall, "unit": "trajectory steps/s", "timing": "host wall clock around the C-ABI call incl. copies"})
        out["rollout_batches"] = rec
    except Exception as exc:
        out["rollout_batches"] = {"error": f"{type(exc).__name__}: {exc}"}
    try:
        from pyro_b200 import dynamicprogramming
        from tests.cases import CASES, build_case
        golden = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        # ---- spline class: parity on the fixture, time per backup on the cfg2 grid ---------------------------------------
        gold = golden("spline_pend_51x51x11")
        _, grid, cf = build_case(CASES["pend_51x51x11"])
        dp = dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid, cf)
        dp.verbose = False
        k = int(gold["snapshots"][1])
        dp.compute_steps(k)
        J_ref = gold[f"J_{k}"]
        rec = {"parity": {"against": "tests/golden/spline_pend_51x51x11.npz: DynamicProgramming2DRectBivariateSpline of the unmodified reference",
                          "sweeps": k, "J_Linf_error": float(np.abs(dp.J - J_ref).max() / np.abs(J_ref).max()),
                          "pi_mismatches": int((dp.pi != gold[f"pi_{k}"]).sum())}}
        case = spline_case or WORKLOADS["cfg2"]
        _, grid, cf = build_case(case)
        dp = dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid, cf)
        eng, K = dp._engine, 5
        eng.sweep(2)
        eng.sweep(K)
        ms = eng.last_sweep_ms / K
        rec.update({"workload": f"SinglePendulum {case['x_grid_dim']} x {case['u_grid_dim']} in table mode (tables from pdp_build_tables), spline refitted every backup",
                    "ms_per_backup": ms, "value": float(grid.nodes_n) * grid.actions_n / ms * 1e3, "unit": "evals/s", "kernel": eng.kernel_info,
                    "timing": "CUDA events inside pdp_sweep: two fit kernels + the sweep kernel per backup"})
        out["spline_class"] = rec
    except Exception as exc:
        out["spline_class"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--clock-sampler":
        return clock_sampler_main(int(sys.argv[2]), float(sys.argv[3]) if len(sys.argv) > 3 else 0.002)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=list(WORKLOADS))
    ap.add_argument("--sub-workloads", default="cfg2,cfg3,cfg4", help="N=1 only: further BASELINE configs reported as sub-records")
    ap.add_argument("--sub-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-after", action="store_true", help="skip the rollout / spline records (SURVEY.md 8f rank 4)")
    ap.add_argument("--no-extra", action="store_true", help="reference arm: skip the cfg1 LUT / base-class measurements")
    args = ap.parse_args()
    wl_name = f"{WORKLOAD_NAMES[args.workload]} (BASELINE {args.workload})"

    if args.impl == "reference":
        run_reference_arm(args, WORKLOADS[args.workload], args.workload, wl_name)
        return
    args.warmup = max(args.warmup, 3)

    ctx = Ctx()
    line = run_workload(ctx, args, args.workload, primary=True)
    subs = {}
    if ctx.world == 1 and args.sub_workloads:
        for key in [k for k in args.sub_workloads.split(",") if k and k != args.workload]:
            try:
                r = run_workload(ctx, args, key, primary=False)
                subs[key] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e", "gpu_launches", "roofline",
                                               "J_Linf_error", "pi_mismatches", "parity", "clocks", "step_ms_min_max")}
            except Exception as exc:   # a sub-record must never cost the headline line
                subs[key] = {"error": f"{type(exc).__name__}: {exc}"}
    if ctx.rank == 0:
        line["sub_records"] = subs
        if ctx.world == 1 and not args.no_after:
            line["after_the_sweep"] = after_the_sweep_records()
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
