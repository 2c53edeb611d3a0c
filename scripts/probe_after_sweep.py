"""Development probe (not the bench), torch-free: timings of the two "after the sweep" additions of round 2 on one GPU.

  * closed-loop Euler rollout batches (pdp_rollout): trajectories x steps per second, wall clock around the C-ABI call
    (H2D of the initial states, the kernel, D2H of every 100th point);
  * the bicubic-spline table sweep (pdp_set_interpolant): ms per backup = two fit kernels + sweep_lut_spline_kernel, CUDA
    events inside pdp_sweep, next to the linear table sweep of the same handle.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pyro_b200 import dynamicprogramming
from tests.cases import CASES, build_case


def rollouts(name, case, sweeps, B, npts, tf):
    sys_, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
    dp.verbose = False
    dp.compute_steps(sweeps)
    rng = np.random.default_rng(1)
    x0 = rng.uniform(np.asarray(sys_.x_lb) * 0.9, np.asarray(sys_.x_ub) * 0.9, (B, sys_.n))
    dp.compute_closed_loop_trajectories(x0[:1024], tf, npts, stride=100)          # warm-up
    t0 = time.perf_counter()
    t, x, u = dp.compute_closed_loop_trajectories(x0, tf, npts, stride=100)
    wall = time.perf_counter() - t0
    inside = np.all((x[:, -1] >= sys_.x_lb) & (x[:, -1] <= sys_.x_ub), axis=1).mean()
    print(json.dumps({"probe": "rollout", "case": name, "dims": case["x_grid_dim"], "policy_sweeps": sweeps, "trajectories": B,
                      "points": npts, "wall_s": round(wall, 4), "steps_per_s": B * (npts - 1) / wall,
                      "final_states_inside_box": float(inside), "finite": bool(np.isfinite(x).all())}), flush=True)


def spline(dims, udims, K=20):
    case = dict(system="SinglePendulum", x_grid_dim=dims, u_grid_dim=udims, xbar=[-3.14, 0.0], INF=300.0)
    _, grid, cf = build_case(case)
    t0 = time.perf_counter()
    dp = dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid, cf)
    setup = time.perf_counter() - t0
    eng = dp._engine
    evals = float(grid.nodes_n) * grid.actions_n
    out = {"probe": "spline_table_sweep", "dims": dims, "udims": udims, "setup_s": round(setup, 2)}
    for which in ("spline3", "linear"):
        eng.set_interpolant(which)
        eng.set_J(np.zeros(grid.nodes_n))
        eng.sweep(3)
        st = eng.sweep(K)
        out[which] = {"kernel": eng.kernel_info, "ms_per_backup": round(eng.last_sweep_ms / K, 4),
                      "evals_per_s": evals / (eng.last_sweep_ms / K) * 1e3, "J_max": float(st[-1][0])}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    jobs = {
        "rollout_pend": lambda: rollouts("pend_201", dict(CASES["pend_51x51x11"], x_grid_dim=[201, 201], u_grid_dim=[21]), 100, 131072, 1001, 10.0),
        "rollout_cartpole": lambda: rollouts("cartpole_41", dict(CASES["cartpole_swingup"], x_grid_dim=[41] * 4, u_grid_dim=[11]), 10, 131072, 501, 5.0),
        "spline501": lambda: spline([501, 501], [51]),
        "spline1001": lambda: spline([1001, 1001], [201], K=5),
    }
    for name in (sys.argv[1:] or list(jobs)):
        try:
            jobs[name]()
        except Exception as e:   # keep going: every line is its own record
            print(json.dumps({"job": name, "error": repr(e)}), flush=True)
