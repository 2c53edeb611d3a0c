"""CPU campaign (no GPU): whole axis-0 planes of a BASELINE configuration AT FULL SIZE through the shipped sweep kernel under
the emulator (tests/emu), every node compared with the C oracle on rough J.

    python scripts/emulated_fullsize_check.py cfg3 all            # every plane (about half a minute per plane of 1e6 nodes)
    python scripts/emulated_fullsize_check.py cfg5 0,1,57,100,200 # chosen planes (5 min each at 201^3 nodes x 961 actions)

One JSON line per plane: nodes, evals, mismatching J / pi entries (expected 0), fraction of nodes with a finite optimum."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import c_oracle
from pyro_b200 import problem
from tests.cases import build_case
from tests.emu import emu

PI = float(np.pi)
WORKLOADS = {   # bench.WORKLOADS without importing torch
    "cfg2": dict(system="SinglePendulum", x_grid_dim=[1001, 1001], u_grid_dim=[201], xbar=[-3.14, 0.0], INF=300.0, dt=0.05),
    "cfg3": dict(system="TwoLinkManipulator", x_grid_dim=[101] * 4, u_grid_dim=[21, 21], INF=1000.0, dt=0.05),
    "cfg4": dict(system="CartPole", x_grid_dim=[151] * 4, u_grid_dim=[51], xbar=[0.0, PI, 0.0, 0.0], INF=1000.0, dt=0.05),
    "cfg5": dict(system="DoublePendulum", x_grid_dim=[201] * 4, u_grid_dim=[31, 31], dt=0.1,
                 x_lb=[-5.0, -1.5, -4.0, -4.0], x_ub=[0.5, 4.0, 5.5, 7.0], u_lb=[-12.0, -12.0], u_ub=[12.0, 12.0],
                 xbar=[0.0, 0.0, 0.0, 0.0], Q=[1.0, 0.5, 0.1, 0.05], R=[0.05, 0.05], INF=1000.0, EPS=1.0),
}


def main():
    name, spec = sys.argv[1], sys.argv[2]
    case = WORKLOADS[name]
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    n0 = P.dims[0]
    plane = P.N // n0
    planes = list(range(n0)) if spec == "all" else [int(p) for p in spec.split(",")]
    rng = np.random.default_rng(5)
    J0 = np.empty(P.N)
    for i in range(n0):                       # rough J, drawn plane by plane (memory)
        J0[i * plane:(i + 1) * plane] = rng.uniform(0, case["INF"], plane)
    J, pi = np.zeros(P.N), np.zeros(P.N, dtype=np.int64)
    bad = 0
    for p in planes:
        t = time.time()
        emu.sweep_planes(P, J0, J, pi, p, p + 1, lanes=1)
        te = time.time() - t
        Jr, pr = c_oracle.sweep_fused(P, J0, p * plane, (p + 1) * plane)
        sl = slice(p * plane, (p + 1) * plane)
        rec = {"workload": name, "plane": p, "nodes": plane, "evals": plane * P.A, "emulator_s": round(te, 1),
               "J_mismatches": int((J[sl] != Jr).sum()), "pi_mismatches": int((pi[sl] != pr).sum()),
               "finite_fraction": float((Jr < case["INF"]).mean())}
        bad += rec["J_mismatches"] + rec["pi_mismatches"]
        print(json.dumps(rec), flush=True)
    print(json.dumps({"workload": name, "planes_checked": len(planes), "of": n0, "nodes_checked": len(planes) * plane,
                      "mismatches": bad}), flush=True)


if __name__ == "__main__":
    main()
