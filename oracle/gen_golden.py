"""Generate tests/golden/*.npz by running the UNMODIFIED reference (tier-0) in this container.

    PYTHONPATH=. python oracle/gen_golden.py            # needs /root/reference (or $PYRO_REF)

The reference ships no tests or golden vectors for this path, so these
fixtures — outputs of the reference's own DynamicProgrammingWithLookUpTable
(pyro/planning/dynamicprogramming.py:505-570) on small grids of the four BASELINE systems —
are the pin for the oracle and for the CUDA path.  Each file holds the case definition (JSON),
J and pi snapshots after the listed sweep counts, and a strided sample of the reference's
x_next_table / G tables (discretizer.py:349, dynamicprogramming.py:523).
The same case definitions are replayed by tests/cases.py on the mirrors of pyro_b200.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from tests.cases import CASES, POLICY_CASES, LinearFeedback  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


build_reference = ref_loader.build_reference


def main():
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        with ref_loader.quiet():
            sys_, grid, cf, dp = build_reference(ns, case)
            out = {"case": json.dumps(case), "J0": dp.J.copy()}
            stride = case.get("table_stride", 1)
            out["table_stride"] = stride
            out["x_next_sample"] = grid.x_next_table[::stride].copy()
            out["x_next_isok_sample"] = grid.x_next_isok[::stride].copy()
            out["action_isok_sample"] = grid.action_isok[::stride].copy()
            out["G_sample"] = dp.G[::stride].copy()
            k = 0
            for target in case["snapshots"]:
                dp.compute_steps(target - k)
                k = target
                out[f"J_{k}"] = dp.J.copy()
                out[f"pi_{k}"] = dp.pi.astype(np.int64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: N={grid.nodes_n} A={grid.actions_n} snapshots={case['snapshots']} "
              f"J_max={dp.J.max():.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main_policy():
    """Fixtures of the reference's PolicyEvaluatorWithLookUpTable (dynamicprogramming.py:677-752)."""
    ns = ref_loader.load()
    for name, case in POLICY_CASES.items():
        with ref_loader.quiet():
            sys_, grid, cf, _ = build_reference(ns, case)
            ctl = LinearFeedback(**case["ctl"])
            pe = ns.dynamicprogramming.PolicyEvaluatorWithLookUpTable(ctl, grid, cf)
            pe.alpha = case.get("alpha", 1.0)
            out = {"case": json.dumps(case), "J0": pe.J.copy(), "x_next_table": pe.x_next_table.copy(), "G": pe.G.copy()}
            k = 0
            for target in case["snapshots"]:
                pe.compute_steps(target - k)
                k = target
                out[f"J_{k}"] = pe.J.copy()
            # the base class (per-node Python loop, :636-672): exact INF wherever the input is disallowed
            pb = ns.dynamicprogramming.PolicyEvaluator(ctl, grid, cf)
            pb.alpha = case.get("alpha", 1.0)
            kb = case["snapshots"][1]
            pb.compute_steps(kb)
            out[f"Jbase_{kb}"] = pb.J.copy()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: N={grid.nodes_n} snapshots={case['snapshots']} J_max={pe.J.max():.6f} "
              f"INF nodes={int((pe.G == cf.INF).sum())} -> {os.path.getsize(path) / 1024:.0f} KiB")


ROLLOUT_CASES = {   # golden case -> (sweeps of the policy used, tf, npts, initial states)
    "pend_51x51x11": (200, 2.0, 401, [[-3.14, 0.0], [-2.0, 1.0], [0.5, -0.5], [3.0, 2.0], [-6.2, 0.3], [1.0, 6.0]]),
    "cartpole_swingup": (20, 0.5, 101, [[0.0, 0.1, 0.0, 0.0], [1.0, 3.0, -1.0, 0.5], [-2.0, -3.0, 2.0, -2.0], [5.0, 1.0, 4.0, 4.0]]),
    "twolink_soft": (30, 0.5, 101, [[0.1, 0.2, 0.0, 0.0], [1.0, -1.0, 0.5, -0.5], [-2.0, 2.5, -1.5, 2.0]]),
    "dpend_example": (6, 0.5, 101, [[-3.14, 0.0, 0.0, 0.0], [0.5, 0.5, 1.0, -1.0], [-1.0, 1.5, -2.0, 2.0]]),
}


def main_rollouts():
    """Fixtures of the reference's closed-loop 'euler' simulation under its LookUpTableController
    (simulation.py:298-324 via ClosedLoopSystem.compute_trajectory controller.py:517-530; dynamicprogramming.py:27-107):
    policy = pi of the committed value-iteration fixture after the given number of sweeps."""
    ns = ref_loader.load()
    for name, (k, tf, npts, x0s) in ROLLOUT_CASES.items():
        case = CASES[name]
        gold = np.load(os.path.join(OUT, name + ".npz"))
        with ref_loader.quiet():
            sys_, grid, cf, dp = build_reference(ns, case)
            dp.compute_steps(k)
            if f"pi_{k}" in gold:
                assert np.array_equal(dp.pi, gold[f"pi_{k}"])
            ctl = dp.get_lookup_table_controller()
            cl_sys = ctl + sys_
            xs, us = [], []
            for x0 in x0s:
                cl_sys.x0 = np.array(x0, dtype=float)
                traj = cl_sys.compute_trajectory(tf, npts, 'euler')
                xs.append(traj.x.copy())
                us.append(traj.u.copy())
        path = os.path.join(OUT, "rollout_" + name + ".npz")
        np.savez_compressed(path, sweeps=k, tf=tf, npts=npts, x0=np.array(x0s, float), pi=dp.pi.astype(np.int64),
                            x=np.array(xs), u=np.array(us))
        print(f"rollout_{name}: {len(x0s)} trajectories x {npts} points -> {os.path.getsize(path) / 1024:.0f} KiB")


SPLINE_CASES = {   # golden case (2-D) -> sweep counts of the snapshots
    "pend_51x51x11": [1, 5, 20],
    "pend_time_41x61x7": [1, 4, 12],
}


def main_spline():
    """Fixtures of the reference's DynamicProgramming2DRectBivariateSpline (dynamicprogramming.py:578-614) on the 2-D
    value-iteration cases: J / pi snapshots, plus for the last one the reference's Q gap between its best and second
    best action per node (where it is ~0 the argmin is decided by rounding and a floating-point-parity path may differ)."""
    ns = ref_loader.load()
    for name, snaps in SPLINE_CASES.items():
        case = CASES[name]
        with ref_loader.quiet():
            sys_, grid, cf, dp0 = build_reference(ns, case)
            dp = ns.dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid, cf)
            dp.alpha = case.get("alpha", 1.0)
            out = {"J0": dp.J.copy()}
            k = 0
            for target in snaps:
                dp.compute_steps(target - k)
                k = target
                out[f"J_{k}"] = dp.J.copy()
                out[f"pi_{k}"] = dp.pi.astype(np.int64)
                Qs = np.sort(dp.Q, axis=1)
                out[f"gap_{k}"] = Qs[:, 1] - Qs[:, 0]
        path = os.path.join(OUT, "spline_" + name + ".npz")
        np.savez_compressed(path, snapshots=np.array(snaps), **out)
        print(f"spline_{name}: snapshots={snaps} J_max={dp.J.max():.6f} J_min={dp.J.min():.6f} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main_3d_example():
    """Fixture of a 3-D example of the reference (helicopter_tunnel.py, obstacles in isavalidstate) on a coarse grid: the
    reference's own tables (x_next_table, G) and J / pi snapshots of DynamicProgrammingWithLookUpTable, alpha = 0.999."""
    ns = ref_loader.load()
    from pyro.dynamic import drone
    from tests.cases import helicopter_tunnel_example
    snaps = [1, 5, 25]
    with ref_loader.quiet():
        sys_, grid, qcf = helicopter_tunnel_example(drone, ns.costfunction, ns.discretizer)
        dp = ns.dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, qcf)
        dp.alpha = 0.999
        out = {"J0": dp.J.copy(), "x_next_table": grid.x_next_table.copy(), "G": dp.G.copy(), "x_next_isok": grid.x_next_isok.copy()}
        k = 0
        for target in snaps:
            dp.compute_steps(target - k)
            k = target
            out[f"J_{k}"] = dp.J.copy()
            out[f"pi_{k}"] = dp.pi.astype(np.int64)
    path = os.path.join(OUT, "helicopter_tunnel_15x13x11.npz")
    np.savez_compressed(path, snapshots=np.array(snaps), x_grid_dim=np.array(grid.x_grid_dim), u_grid_dim=np.array(grid.u_grid_dim),
                        alpha=0.999, **out)
    print(f"helicopter_tunnel: N={grid.nodes_n} A={grid.actions_n} invalid arrivals {1 - grid.x_next_isok.mean():.3f} "
          f"J_max={dp.J.max():.3f} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    if "--rollouts-only" in sys.argv:
        main_rollouts()
        sys.exit(0)
    if "--spline-only" in sys.argv:
        main_spline()
        sys.exit(0)
    if "--3d-only" in sys.argv:
        main_3d_example()
        sys.exit(0)
    if "--policy-only" not in sys.argv:
        main()
    main_policy()
    main_rollouts()
    main_spline()
    main_3d_example()
