"""Development probe: LUT-mode sweep with ONE column per node (policy evaluation, dynamicprogramming.py:743-752) on a
large 2-D / 4-D grid: the HBM-bound member of the family.  Algorithmic bytes per node: x_next (8n) + G (8) + J_next of the
node (8, for the fused dJ) + J (8) + pi (8); the 2^n corner gathers are served by L1/L2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyro_b200 import systems, discretizer, costfunction, problem
from pyro_b200.engine import Engine

def run(kind, dims, A=1):
    s = systems.SYSTEMS[kind]()
    g = discretizer.GridDynamicSystem(s, dims, [3] * s.m)
    cf = costfunction.QuadraticCostFunction.from_sys(s)
    P = problem.extract(g, cf, 1.0, lut_actions=A)
    N, n = P.N, P.n
    rng = np.random.default_rng(0)
    X = g.state_from_node_id                       # (N, n)
    step = np.array([(s.x_ub[i] - s.x_lb[i]) / (dims[i] - 1) for i in range(n)])
    x_next = X[:, None, :] + rng.uniform(-2.5, 2.5, (N, A, n)) * step   # a few cells away, some out of the box
    G = rng.uniform(0, 1, (N, A))
    eng = Engine(P)
    eng.set_lut(x_next, G)
    eng.set_J(rng.uniform(0, 300, N))
    eng.sweep(3)
    K = 10
    eng.sweep(K)
    ms = eng.last_sweep_ms / K
    bytes_node = A * (8 * n + 8) + 24
    print(json.dumps({"case": f"{kind} {dims} LUT A={A}", "nodes": N, "ms_per_sweep": round(ms, 4), "nodes_per_s": N / ms * 1e3,
                      "algorithmic_GBs": N * bytes_node / ms / 1e6, "bytes_per_node": bytes_node}), flush=True)
    eng.close()

if __name__ == "__main__":
    run("SinglePendulum", [4001, 4001])
    run("CartPole", [71, 71, 71, 71])
    run("SinglePendulum", [2001, 2001], A=8)
