"""Development probe (not the bench): time sweeps of a few configurations on one GPU.

    python scripts/probe_perf.py cfg3 cfg4 dp81          # names: bench.WORKLOADS + the reduced grids below
    PYRODP_MECH2=generic python scripts/probe_perf.py cfg3   # A/B of the 4-D kernel variants (generic|direct|cache)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bench import WORKLOADS
from pyro_b200 import problem
from pyro_b200.engine import Engine
from tests.cases import build_case

CASES = dict(WORKLOADS)
CASES.update({
    "tl61": dict(WORKLOADS["cfg3"], x_grid_dim=[61] * 4),
    "cp101": dict(WORKLOADS["cfg4"], x_grid_dim=[101] * 4),
    "cp61": dict(WORKLOADS["cfg4"], x_grid_dim=[61] * 4),
    "dp61": dict(WORKLOADS["cfg5"], x_grid_dim=[61] * 4),
    "dp81": dict(WORKLOADS["cfg5"], x_grid_dim=[81] * 4),
    "dp101": dict(WORKLOADS["cfg5"], x_grid_dim=[101] * 4),
})


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "tl61", "cp101", "dp61"]
    for name in names:
        case = CASES[name]
        _, g, cf = build_case(case)
        t0 = time.time()
        eng = Engine(problem.extract(g, cf, 1.0))
        eng.eval_terminal_cost()
        evals = float(g.nodes_n) * g.actions_n
        eng.sweep(3 if evals < 1e11 else 1)
        K = 10 if evals < 5e9 else (3 if evals < 1e11 else 1)
        st = eng.sweep(K)
        ms = eng.last_sweep_ms / K
        print(json.dumps({"case": name, "sys": case["system"], "dims": case["x_grid_dim"], "udims": case["u_grid_dim"],
                          "kernel": eng.kernel_info, "ms_per_sweep": round(ms, 4), "evals_per_s": evals / ms * 1e3,
                          "jmax": st[-1][0], "dmax": st[-1][1], "dmin": st[-1][2], "setup_s": round(time.time() - t0, 2)}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
