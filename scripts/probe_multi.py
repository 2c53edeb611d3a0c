"""Development probe: the single-process multi-GPU engine (pdp_multi_*) on every visible GPU — ms per sweep and a sampled
check against the C oracle.  Usage: python scripts/probe_multi.py cfg4 [n_parts]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bench import WORKLOADS
from oracle import c_oracle
from pyro_b200 import problem
from pyro_b200.engine import MultiEngine, device_count
from tests.cases import build_case


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    n_parts = int(sys.argv[2]) if len(sys.argv) > 2 else device_count()
    _, g, cf = build_case(WORKLOADS[name])
    P = problem.extract(g, cf, 1.0)
    t0 = time.time()
    eng = MultiEngine(P, n_parts=n_parts)
    eng.eval_terminal_cost()
    eng.sweep(2)
    K = 5
    t1 = time.perf_counter()
    eng.sweep(K)
    ms = 1e3 * (time.perf_counter() - t1) / K
    J_next = eng.get_J_next()
    J, pi = eng.get_J(), eng.get_pi()
    rng = np.random.default_rng(3)
    plane = P.N // P.dims[0]
    starts = [int(s) for s in rng.integers(0, P.N - 256, 12)] + [plane * (r * P.dims[0] // n_parts) - 128 for r in range(1, n_parts)]
    bad = 0
    for lo in starts:
        Jr, pr = c_oracle.sweep_fused(P, J_next, lo, lo + 256)
        bad += int(not (np.array_equal(J[lo:lo + 256], Jr) and np.array_equal(pi[lo:lo + 256], pr)))
    evals = float(P.N) * P.A
    print(json.dumps({"case": name, "engine": "MultiEngine (one process, one host thread)", "n_parts": eng.n_parts, "devices": eng.devices,
                      "kernel": eng.kernel_info, "ms_per_sweep_wall": round(ms, 3), "evals_per_s": evals / ms * 1e3,
                      "ranges_checked": len(starts), "ranges_mismatching": bad, "setup_s": round(time.time() - t0, 1)}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
