"""Development probe (not the bench), torch-free: the pendulum sweep at BASELINE cfg2 (1001^2 x 201) with both MONO loops
of sweep_pendulum_kernel — ms per sweep (CUDA events inside pdp_sweep, J resident, no L2 flush: use for A/B only) and
bit equality of J / pi between the loops after the timed sweeps.

    python scripts/probe_pend.py                 # loop nest (PYRODP_PEND_LOOP=2) and the shipped pair loop, same library
    PYRODP_LIB=pyro_b200/libpyrodp_b9.so python scripts/probe_pend.py nest     # another build of the library
    python scripts/probe_pend.py nest --sweeps 6 # short run for ncu
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pyro_b200 import problem
from pyro_b200.engine import Engine
from tests.cases import build_case

CFG2 = dict(system="SinglePendulum", x_grid_dim=[1001, 1001], u_grid_dim=[201], xbar=[-3.14, 0.0], INF=300.0)
LOOPS = {"nest": "2", "pair": "1"}


def main():
    sweeps = int(sys.argv[sys.argv.index("--sweeps") + 1]) if "--sweeps" in sys.argv else 200
    args = [a for a in sys.argv[1:] if a in LOOPS]
    _, g, cf = build_case(CFG2)
    evals = float(g.nodes_n) * g.actions_n
    results = {}
    for name in args or ["nest", "pair", "nest"]:
        os.environ["PYRODP_PEND_LOOP"] = LOOPS[name]
        eng = Engine(problem.extract(g, cf, 1.0))
        eng.eval_terminal_cost()
        eng.sweep(5)
        eng.sweep(sweeps)
        ms = eng.last_sweep_ms / sweeps
        J, pi = eng.get_J(), eng.get_pi()
        same = None
        if results and name not in results:
            J0, pi0 = next(iter(results.values()))
            same = bool(np.array_equal(J, J0) and np.array_equal(pi, pi0))
        results.setdefault(name, (J, pi))
        print(json.dumps({"case": "cfg2", "loop": name, "lib": os.environ.get("PYRODP_LIB", "libpyrodp.so"), "kernel": eng.kernel_info,
                          "sweeps": sweeps, "ms_per_sweep": round(ms, 5), "evals_per_s": evals / ms * 1e3,
                          "bit_equal_to_first": same, "J_max": float(J.max())}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
