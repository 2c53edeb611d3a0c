"""Development probe (not the bench): time sweeps of a few configurations on one GPU."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyro_b200 import systems, costfunction, discretizer, problem
from pyro_b200.engine import Engine

CASES = {
    "cfg1": ("SinglePendulum", [51, 51], [11], [-3.14, 0.0], 300.0, 0.05),
    "cfg2": ("SinglePendulum", [1001, 1001], [201], [-3.14, 0.0], 300.0, 0.05),
    "tl61": ("TwoLinkManipulator", [61] * 4, [21, 21], None, 1000.0, 0.05),
    "cp101": ("CartPole", [101] * 4, [51], [0, np.pi, 0, 0], 1000.0, 0.05),
    "dp61": ("DoublePendulum", [61] * 4, [31, 31], None, 1000.0, 0.05),
    "cfg3": ("TwoLinkManipulator", [101] * 4, [21, 21], None, 1000.0, 0.05),
    "cfg4": ("CartPole", [151] * 4, [51], [0, np.pi, 0, 0], 1000.0, 0.05),
    "cfg5": ("DoublePendulum", [201] * 4, [31, 31], None, 1000.0, 0.05),
    "cp61": ("CartPole", [61] * 4, [51], [0, np.pi, 0, 0], 1000.0, 0.05),
}

def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "tl61", "cp101", "dp61"]
    for name in names:
        kind, xd, ud, xbar, INF, dt = CASES[name]
        s = systems.SYSTEMS[kind]()
        g = discretizer.GridDynamicSystem(s, xd, ud, dt)
        cf = costfunction.QuadraticCostFunction.from_sys(s)
        if xbar is not None:
            cf.xbar = np.array(xbar, float)
        cf.INF = INF
        t0 = time.time()
        eng = Engine(problem.extract(g, cf, 1.0))
        eng.eval_terminal_cost()
        evals_ = g.nodes_n * g.actions_n
        eng.sweep(3 if evals_ < 1e11 else 1)
        K = 10 if evals_ < 5e9 else (3 if evals_ < 1e11 else 1)
        st = eng.sweep(K)
        ms = eng.last_sweep_ms / K
        evals = g.nodes_n * g.actions_n
        print(json.dumps({"case": name, "sys": kind, "dims": xd, "udims": ud, "ms_per_sweep": round(ms, 4),
                          "evals_per_s": evals / ms * 1e3, "jmax": st[-1][0], "dmax": st[-1][1], "dmin": st[-1][2],
                          "setup_s": round(time.time() - t0, 2)}), flush=True)
        eng.close()

if __name__ == "__main__":
    main()
