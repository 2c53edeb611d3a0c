"""Randomised parity (seeded, deterministic): random plants / bounds / grids / action ladders / costs / discounts.

  * the product's kernels (compiled for the CPU emulator) against the C oracle: every lane split, both MONO loops of the
    pendulum kernel, the order-agnostic and the range-skipping 4-D kernels, the fused dJ statistics — bit for bit;
  * the C oracle against the LIVE unmodified reference (where it is importable: this container) after 1-3 sweeps of its
    DynamicProgrammingWithLookUpTable — bit for bit.  This widens the pin of the oracle beyond the committed fixtures.

`python tests/test_fuzz.py kernels|reference|halo|tables|rollouts|spline|lut <first seed> <count>` runs longer campaigns (DESIGN.md section 3 records one:
3000 + 3000 cases, no mismatch)."""
import sys

import numpy as np
import pytest

SYSTEMS = ["SinglePendulum", "SinglePendulum", "CartPole", "TwoLinkManipulator", "DoublePendulum"]


def random_case(rng, tiny=False):
    kind = str(rng.choice(SYSTEMS))
    case = dict(system=kind, INF=float(rng.choice([300.0, 1000.0, 50.0])))
    if kind == "SinglePendulum":
        n, m = 2, 1
        case["x_grid_dim"] = [int(rng.integers(2, 14 if tiny else 40)), int(rng.integers(2, 30 if tiny else 200))]
        case["u_grid_dim"] = [int(rng.integers(1, 12 if tiny else 70))]
        if rng.random() < 0.5:
            case["sys_params"] = {"d1": float(rng.uniform(0, 1))}
    else:
        n, m = 4, (1 if kind == "CartPole" else 2)
        hi = (5, 5, 6, 7) if tiny else (8, 8, 14, 20)
        case["x_grid_dim"] = [int(rng.integers(2, h)) for h in hi]
        case["u_grid_dim"] = [int(rng.integers(1, 5 if tiny else 12)) for _ in range(m)]
    lo, hi = -rng.uniform(0.3, 7.0, n), rng.uniform(0.3, 7.0, n)
    if rng.random() < 0.3:
        shift = rng.uniform(-2, 2, n)
        lo, hi = lo + shift, hi + shift
    case["x_lb"], case["x_ub"] = lo.tolist(), hi.tolist()
    ul = rng.uniform(0.5, 15.0, m)
    case["u_lb"], case["u_ub"] = (-ul).tolist(), (ul * rng.uniform(0.5, 1.0, m)).tolist()
    case["dt"] = float(rng.choice([0.05, 0.1, 0.02, 0.2]))
    case["xbar"] = rng.uniform(lo, hi).tolist()
    case["cost"] = str(rng.choice(["quadratic", "quadratic", "time", "reach", "domaincheck"]))
    if case["cost"] in ("quadratic", "domaincheck"):
        case["Q"], case["R"] = rng.uniform(0, 3, n).tolist(), rng.uniform(0.01, 2, m).tolist()
    case["EPS"] = float(rng.choice([1e-3, 0.5, 1.0]))
    case["alpha"] = float(rng.choice([1.0, 1.0, 0.9, 0.5]))
    return case


def kernels_vs_oracle(seed):
    """Mismatch descriptions of one random case (empty list = parity)."""
    from oracle import c_oracle
    from pyro_b200 import problem
    from tests.cases import build_case
    from tests.emu import emu
    rng = np.random.default_rng(seed)
    case = random_case(rng)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case["alpha"])
    assert P.system_id != 0, case
    J0 = rng.uniform(0, case["INF"], P.N) if rng.random() < 0.8 else c_oracle.terminal(P)
    Jr, pr = c_oracle.sweep_fused(P, J0)
    variants = [(1, False, None)] + ([(4, False, None)] if P.A >= 4 else []) + ([(16, False, None)] if P.A >= 16 else [])
    if case["system"] == "SinglePendulum":
        variants += [(1, 2, None), (1, True, None)] + ([(4, 2, None)] if P.A >= 4 else [])   # loop nest, order-agnostic loop
    else:
        variants += [(1, False, "generic")]
    bad = []
    for lanes, loop, mech2 in variants:
        J, pi, st = emu.sweep(P, J0, lanes=lanes, force_generic=loop, mech2=mech2)
        d = J - J0
        if not (np.array_equal(J, Jr) and np.array_equal(pi, pr) and st[0] == J.max() and st[1] == d.max() and st[2] == d.min()):
            bad.append((seed, lanes, loop, mech2, int((J != Jr).sum()), int((pi != pr).sum()), case))
    return bad


def oracle_vs_reference(ns, seed):
    from oracle import c_oracle, ref_loader
    from pyro_b200 import problem
    from tests.cases import build_case
    rng = np.random.default_rng(seed)
    case = random_case(rng, tiny=True)
    k = int(rng.integers(1, 4))
    with ref_loader.quiet():
        _, _, _, rdp = ref_loader.build_reference(ns, case)
        rdp.compute_steps(k)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case["alpha"])
    assert P.system_id != 0, case
    J, pi, _ = c_oracle.run(P, k)
    if np.array_equal(J, rdp.J) and np.array_equal(pi, rdp.pi):
        return []
    return [(seed, k, int((J != rdp.J).sum()), int((pi != rdp.pi).sum()), case)]


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
def test_random_problems_emulated_kernels_equal_c_oracle():
    bad = [b for seed in range(100, 220) for b in kernels_vs_oracle(seed)]
    assert not bad, bad[:3]


@pytest.mark.filterwarnings("ignore::RuntimeWarning", "ignore::DeprecationWarning")
def test_random_problems_c_oracle_equals_the_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("unmodified reference not present")
    ns = ref_loader.load()
    bad = [b for seed in range(7000, 7150) for b in oracle_vs_reference(ns, seed)]
    assert not bad, bad[:3]


def halo_sufficiency(seed):
    """One random case cut into random slabs over axis 0: every slab's launch sees J_next only on slab + the halo that
    pdp_compute_halo promises (NaN elsewhere) and must reproduce the whole-grid backup."""
    import ctypes as C
    from pyro_b200 import _lib, problem
    from tests.cases import build_case
    from tests.emu import emu
    rng = np.random.default_rng(seed)
    case = random_case(rng)
    n0 = int(rng.integers(6, 40))
    case["x_grid_dim"] = [n0] + ([int(rng.integers(2, 60))] if case["system"] == "SinglePendulum"
                                 else [int(rng.integers(2, 5)), int(rng.integers(2, 8)), int(rng.integers(2, 10))])
    if rng.random() < 0.5:          # velocity bounds small against the position range: thin halos, many slabs
        k = 1 if case["system"] == "SinglePendulum" else 2
        case["x_lb"][k], case["x_ub"][k] = -float(rng.uniform(0.2, 1.5)), float(rng.uniform(0.2, 1.5))
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case["alpha"])
    lo, hi = C.c_int32(), C.c_int32()
    _lib.check(_lib.load().pdp_compute_halo(C.byref(P.c), C.byref(lo), C.byref(hi)))
    lo, hi = lo.value, hi.value
    plane = P.N // n0
    J0 = rng.uniform(0, case["INF"], P.N)
    J_ref, pi_ref, _ = emu.sweep(P, J0, lanes=1)
    cuts = sorted(set([0, n0] + [int(c) for c in rng.integers(1, n0, int(rng.integers(1, 5)))]))
    J, pi = np.full(P.N, -1.0), np.full(P.N, -1, dtype=np.int64)
    for b, e in zip(cuts[:-1], cuts[1:]):
        seen = np.full(P.N, np.nan)
        a0, a1 = max(0, b - lo), min(n0, e + hi)
        seen[a0 * plane:a1 * plane] = J0[a0 * plane:a1 * plane]
        emu.sweep_planes(P, seen, J, pi, b, e, lanes=1)
    if np.array_equal(J, J_ref) and np.array_equal(pi, pi_ref):
        return []
    return [(seed, lo, hi, cuts, int((J != J_ref).sum()), int(np.isnan(J).sum()), case)]


@pytest.mark.filterwarnings("ignore::RuntimeWarning")
def test_random_slab_layouts_need_only_the_promised_halo():
    bad = [b for seed in range(300, 380) for b in halo_sufficiency(seed)]
    assert not bad, bad[:3]



def tables_vs_reference(ns, seed):
    """build_tables_kernel (emulated) against the live reference's x_next_table, x_next_isok and cost table G."""
    from oracle import ref_loader
    from pyro_b200 import problem
    from tests.cases import build_case
    from tests.emu import emu
    rng = np.random.default_rng(seed)
    case = random_case(rng, tiny=True)
    with ref_loader.quiet():
        _, rgrid, _, rdp = ref_loader.build_reference(ns, case)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case["alpha"])
    xn, ok, G = emu.build_tables(P)
    if np.array_equal(xn, rgrid.x_next_table) and np.array_equal(ok, rgrid.x_next_isok) and np.array_equal(G, rdp.G):
        return []
    return [(seed, int((xn != rgrid.x_next_table).sum()), int((ok != rgrid.x_next_isok).sum()), int((G != rdp.G).sum()), case)]


@pytest.mark.filterwarnings("ignore::RuntimeWarning", "ignore::DeprecationWarning")
def test_random_problems_table_builder_equals_the_live_reference_tables():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("unmodified reference not present")
    ns = ref_loader.load()
    bad = [b for seed in range(9000, 9120) for b in tables_vs_reference(ns, seed)]
    assert not bad, bad[:3]



def rollouts_vs_reference(ns, seed, rtol=1e-8):
    """rollout_kernel (emulated) against the live reference's closed-loop 'euler' simulation under its LookUpTableController,
    with a RANDOM policy table (rough input tables: the interpolation conventions matter), random plant parameters,
    initial states inside and outside the grid."""
    from oracle import ref_loader
    from pyro_b200 import problem
    from tests.cases import build_case
    from tests.emu import emu
    rng = np.random.default_rng(seed)
    case = random_case(rng, tiny=True)
    case["x_grid_dim"] = [max(d, 3) for d in case["x_grid_dim"]]
    case["cost"], case["alpha"] = "quadratic", 1.0
    case.setdefault("Q", [1.0] * len(case["x_grid_dim"]))
    case.setdefault("R", [1.0] * len(case["u_grid_dim"]))
    with ref_loader.quiet():
        rsys, rgrid, _, _ = ref_loader.build_reference(ns, case)
        pi = rng.integers(0, rgrid.actions_n, rgrid.nodes_n)
        ctl = ns.dynamicprogramming.LookUpTableController(rgrid, pi)
        cl = ctl + rsys
        npts, tf = int(rng.integers(5, 40)), float(rng.uniform(0.1, 1.0))
        lo, hi = np.asarray(rsys.x_lb), np.asarray(rsys.x_ub)
        x0s = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), (4, rsys.n))
        xs, us = [], []
        for x0 in x0s:
            cl.x0 = x0.copy()
            traj = cl.compute_trajectory(tf, npts, 'euler')
            xs.append(traj.x.copy()); us.append(traj.u.copy())
    xs, us = np.array(xs), np.array(us)
    sys_, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    x, u = emu.rollout(P, pi, problem.plant_parameters(sys_, P.system_id), x0s, npts, tf / (npts - 1))
    finite = np.isfinite(xs).all()
    scale = max(1.0, np.abs(xs[np.isfinite(xs)]).max()) if np.isfinite(xs).any() else 1.0
    if finite and scale < 1e6:      # trajectories that blow up (stiff random plants at large dt) amplify the last bit: not compared
        if np.abs(x - xs).max() > rtol * scale or np.abs(u - us).max() > rtol * max(1.0, np.abs(us).max()):
            return [(seed, float(np.abs(x - xs).max()), float(np.abs(u - us).max()), scale, case)]
    return []


def spline_vs_scipy(seed, rtol=1e-9):
    """Host plan + fit kernels + sweep_lut_spline_kernel (emulated) on random tables and rough J against the oracle's scipy call."""
    from oracle import np_oracle as npo
    from pyro_b200 import problem
    from tests.emu import emu
    rng = np.random.default_rng(seed)
    dims, A = [int(rng.integers(4, 30)), int(rng.integers(4, 30))], int(rng.integers(2, 9))
    lb, ub = -rng.uniform(0.5, 5, 2), rng.uniform(0.5, 5, 2)
    levels = [np.linspace(lb[i], ub[i], dims[i]) for i in range(2)]
    N = dims[0] * dims[1]
    X = np.stack([g.reshape(-1) for g in np.meshgrid(*levels, indexing="ij")], axis=1)
    x_next = X[:, None, :] + rng.normal(0, 0.3 * (ub - lb) / np.array(dims), (N, A, 2)) * rng.choice([1.0, 5.0])
    G = rng.uniform(0, 2, (N, A))
    J0 = rng.uniform(0, 100, N)
    alpha = float(rng.choice([1.0, 0.9]))

    class Sys2:
        n, m = 2, 1
        x_lb, x_ub, u_lb, u_ub = lb, ub, np.array([-1.0]), np.array([1.0])

    class Grid2:
        sys, dt = Sys2(), 0.05
        x_grid_dim, u_grid_dim = np.array(dims), np.array([A])
        x_level, u_level = levels, [np.linspace(-1, 1, A)]

    class Cost2:
        INF = 1000.0
    P = problem.extract(Grid2(), Cost2(), alpha)
    J, pi, _, _ = emu.spline_sweep(P, J0, x_next, G)
    Jr, pr, gap = npo.spline_sweep(levels, dims, J0, x_next, G, alpha)
    scale = np.abs(Jr).max()
    if np.abs(J - Jr).max() > rtol * scale or ((pi != pr) & (gap > 10 * rtol * scale)).any():
        return [(seed, float(np.abs(J - Jr).max()), int((pi != pr).sum()), dims, A)]
    return []


def lut_vs_numpy(seed):
    """sweep_lut_kernel / sweep_policy_kernel (emulated) for n = 2, 3, 4 on random tables against scipy's RGI (bit for bit)."""
    from oracle import np_oracle as npo
    from pyro_b200 import problem
    from tests.emu import emu
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 5))
    dims = [int(rng.integers(2, 9 if n > 2 else 25)) for _ in range(n)]
    A = int(rng.integers(1, 9))
    lb, ub = -rng.uniform(0.5, 5, n), rng.uniform(0.5, 5, n)
    levels = [np.linspace(lb[i], ub[i], dims[i]) for i in range(n)]
    N = int(np.prod(dims))
    X = np.stack([g.reshape(-1) for g in np.meshgrid(*levels, indexing="ij")], axis=1)
    x_next = X[:, None, :] + rng.normal(0, 1.0, (N, A, n)) * (ub - lb) / np.array(dims)
    x_next[::7, 0, :] = X[::7]                  # exact node hits
    x_next[3::11, A - 1, n - 1] = ub[n - 1]     # exactly on the upper bound
    G = np.where(rng.random((N, A)) < 0.1, 500.0, rng.uniform(0, 2, (N, A)))
    J0 = rng.uniform(0, 100, N)
    alpha = float(rng.choice([1.0, 0.97]))

    class SysN:
        m = 1
        x_lb, x_ub, u_lb, u_ub = lb, ub, np.array([-1.0]), np.array([1.0])
    SysN.n = n

    class GridN:
        sys, dt = SysN(), 0.05
        x_grid_dim, u_grid_dim = np.array(dims), np.array([max(A, 1)])
        x_level, u_level = levels, [np.linspace(-1, 1, max(A, 1))]

    class CostN:
        INF = 500.0
    P = problem.extract(GridN(), CostN(), alpha, lut_actions=A if A == 1 else None)
    J, pi, _ = emu.lut_sweep(P, J0, x_next, G)
    Jr, pr = npo.lut_sweep(levels, dims, J0, x_next, G, alpha, use_scipy=True)
    return [] if np.array_equal(J, Jr) and np.array_equal(pi, pr) else [(seed, n, dims, A, int((J != Jr).sum()), int((pi != pr).sum()))]


@pytest.mark.filterwarnings("ignore::RuntimeWarning", "ignore::DeprecationWarning")
def test_random_policies_rollout_kernel_equals_the_live_reference_simulation():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("unmodified reference not present")
    ns = ref_loader.load()
    bad = [b for seed in range(11000, 11060) for b in rollouts_vs_reference(ns, seed)]
    assert not bad, bad[:3]


def test_random_tables_spline_and_lut_kernels_equal_scipy():
    bad = [b for seed in range(12000, 12040) for b in spline_vs_scipy(seed)]
    bad += [b for seed in range(13000, 13080) for b in lut_vs_numpy(seed)]
    assert not bad, bad[:3]


if __name__ == "__main__":
    import os
    import warnings
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    warnings.simplefilter("ignore")
    which, first, count = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    if which in ("reference", "tables", "rollouts"):
        from oracle import ref_loader
        ns = ref_loader.load()
    bad = []
    for seed in range(first, first + count):
        bad += {"kernels": lambda: kernels_vs_oracle(seed), "halo": lambda: halo_sufficiency(seed), "tables": lambda: tables_vs_reference(ns, seed),
                "rollouts": lambda: rollouts_vs_reference(ns, seed), "spline": lambda: spline_vs_scipy(seed), "lut": lambda: lut_vs_numpy(seed),
                "reference": lambda: oracle_vs_reference(ns, seed)}[which]()
    print(f"{which}: seeds {first}..{first + count - 1}: {len(bad)} mismatching variants")
    for b in bad[:10]:
        print(b)
