"""Workload geometry of the 4-D BASELINE configurations (CPU only, a few hundred sampled nodes per configuration):
fraction of (node, action) pairs whose x_next stays in the box, how far x_next[2], x_next[3] move per action step and over
the whole action set (in cells), how often the (c0, c1) base planes change along a warp — and from those the shared-memory
window a 128-node block would have to stage to serve its corner gathers from a TMA tile.

    python scripts/valid_fraction.py > profiles/r02_workload_geometry.txt
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bench import WORKLOADS
from tests.cases import build_case


def main():
    rng = np.random.default_rng(0)
    for name in ["cfg3", "cfg4", "cfg5"]:
        case = WORKLOADS[name]
        s, grid, cf = build_case(case)
        dims, lev, dt = case["x_grid_dim"], grid.x_level, case["dt"]
        U = np.array(list(itertools.product(*grid.u_level)))
        K = 300
        idx = np.stack([rng.integers(0, d, K) for d in dims], 1)
        step = [lev[d][1] - lev[d][0] for d in range(4)]
        valid = tot = posbad = 0
        span2, span3, per_action = [], [], []
        for ix in idx:
            x = np.array([lev[d][ix[d]] for d in range(4)])
            xn = np.array([s.f(x, u) * dt + x for u in U])
            ok = np.all((xn >= s.x_lb) & (xn <= s.x_ub), axis=1)
            valid += ok.sum(); tot += len(U)
            posbad += not (s.x_lb[0] <= xn[0, 0] <= s.x_ub[0] and s.x_lb[1] <= xn[0, 1] <= s.x_ub[1])
            inb = xn[ok]
            if len(inb):
                span2.append((inb[:, 2].max() - inb[:, 2].min()) / step[2]); span3.append((inb[:, 3].max() - inb[:, 3].min()) / step[3])
            per_action.append((np.median(np.abs(np.diff(xn[:, 2]))) / step[2], np.median(np.abs(np.diff(xn[:, 3]))) / step[3]))
        lanes_per_c1 = step[1] / (step[3] * dt)      # x_next[1] = q1 + dq1*dt: cells of axis 1 per lane along axis 3
        lanes_per_c0 = step[0] / (step[2] * dt)
        s2, s3 = float(np.percentile(span2, 90)), float(np.percentile(span3, 90))
        # a block = 128 consecutive nodes of one (i2) row: c1 takes 128/lanes_per_c1 + 2 values, c0 two; rows k2 over the
        # valid actions (+2), columns 128 + span3 + 2
        planes = 2 * (int(128 / lanes_per_c1) + 2)
        win = planes * (int(s2) + 2) * (128 + int(s3) + 2) * 8 / 1024.0
        print(f"{name} {case['system']} {dims} x {case['u_grid_dim']}: in-box pairs {valid / tot:.3f}, nodes with every action out of the box "
              f"(position rows) {posbad / K:.3f}")
        print(f"    per action step: x_next[2] moves {np.mean(per_action, axis=0)[0]:.2f} cells, x_next[3] {np.mean(per_action, axis=0)[1]:.2f} cells; "
              f"over the in-box actions of a node (90th percentile): {s2:.0f} x {s3:.0f} cells")
        print(f"    the (c0,c1) base planes change every {lanes_per_c0:.1f} lanes along axis 2 / every {lanes_per_c1:.1f} lanes along axis 3")
        print(f"    shared-memory window of a 128-node block for TMA-staged corners: {planes} planes x {int(s2) + 2} rows x {128 + int(s3) + 2} "
              f"columns x 8 B = {win:.0f} KB  (227 KB per SM; the kernel needs >= 4 blocks per SM)")


if __name__ == "__main__":
    main()
