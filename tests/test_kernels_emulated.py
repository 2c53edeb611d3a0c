"""The product's fused sweep kernels, compiled from pyro_b200/csrc/*.cuh with g++ and run on the CPU by the
coroutine emulator in tests/emu/ (one coroutine per CUDA thread, real barriers / votes / shuffles), against the
reference fixtures and the C oracle — bit for bit.  This checks the kernels' arithmetic and control flow (cell
walks, lane splits, argmin ties, padded action tables, both pendulum loops) in the CPU suite; the GPU suite
(tests/test_parity_gpu.py) runs the same comparisons on the device through the C ABI."""
import numpy as np
import pytest

from oracle import c_oracle
from pyro_b200 import problem
from tests.cases import CASES, build_case
from tests.conftest import load_golden
from tests.emu import emu


def run_sweeps(P, J, k, lanes, generic=False):
    pi = st = None
    for _ in range(k):
        J_prev = J
        J, pi, st = emu.sweep(P, J_prev, lanes=lanes, force_generic=generic)
        d = J - J_prev
        assert st[0] == J.max() and st[1] == d.max() and st[2] == d.min()   # the fused dJ statistics
    return J, pi


@pytest.mark.parametrize("lanes", [1, 4, 16])
@pytest.mark.parametrize("name", list(CASES))
def test_emulated_kernels_match_reference_goldens(name, lanes):
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    k = case["snapshots"][0] if lanes == 16 else case["snapshots"][1]
    k = min(k, 5)
    J, pi = run_sweeps(P, gold["J0"], k, lanes)
    if f"J_{k}" in gold:
        assert np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"])
    else:   # no snapshot at this sweep count: the C oracle (itself pinned by the fixtures) is the comparison
        Jr, pr, _ = c_oracle.run(P, k)
        assert np.array_equal(J, Jr) and np.array_equal(pi, pr)


@pytest.mark.parametrize("lanes", [1, 4])
@pytest.mark.parametrize("name", ["pend_51x51x11", "pend_time_41x61x7"])
def test_emulated_order_agnostic_pendulum_loop(name, lanes):
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    k = case["snapshots"][0]
    J, pi = run_sweeps(P, gold["J0"], k, lanes, generic=True)
    assert np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"])


@pytest.mark.parametrize("lanes", [1, 4, 16])
@pytest.mark.parametrize("name", ["pend_51x51x11", "pend_time_41x61x7", "pend_reach_41x41x3"])
def test_emulated_loop_nest_of_the_pendulum_kernel(name, lanes):
    """PYRODP_PEND_LOOP=2: the loop-nest variant of the MONO action loop (fewer instructions, same sweep time on a B200;
    the pair loop stays the default and is what every other test runs)."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    k = case["snapshots"][0]
    J, pi = run_sweeps(P, gold["J0"], k, lanes, generic=2)
    assert np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"])


@pytest.mark.parametrize("lanes", [1, 4, 16])
@pytest.mark.parametrize("shape", [
    ([19, 260], [1]),      # one action: the whole loop is padding but the first record
    ([19, 260], [2]),
    ([23, 300], [3]),      # cells much narrower than an action step: every pair skips several cells (general walk)
    ([17, 131], [37]),     # odd number of pairs, lanes past the end of the row in the last block
    ([9, 40], [201]),      # a tenth of a cell per action: the next-cell fast path of the loop nest, 100 pairs per node
    ([5, 3], [7]),         # three levels: two cells
    ([5, 2], [5]),         # two levels: one cell, never a next cell
])
def test_emulated_pendulum_loop_nest_edge_shapes(shape, lanes):
    """Both MONO loops of sweep_pendulum_kernel (pair loop = default, loop nest = PYRODP_PEND_LOOP=2) where their special
    cases live: padding, cell skips, parked lanes (velocity bounds tight enough that many actions leave the box), damping
    (the t[a] table holds B.u - g only)."""
    xd, ud = shape
    for extra in (dict(), dict(sys_params={"d1": 0.3}, x_lb=[-2.0, -1.5], x_ub=[1.0, 2.5]), dict(alpha=0.9, x_lb=[-3.0, -0.4], x_ub=[3.0, 0.4])):
        case = dict(system="SinglePendulum", x_grid_dim=xd, u_grid_dim=ud, xbar=[-3.14, 0.0], INF=300.0, **extra)
        if ud[0] < 4 * lanes and lanes > 1:
            continue   # the library only splits a node over G lanes when A >= 4 (16 for G = 16)
        _, grid, cf = build_case(case)
        P = problem.extract(grid, cf, case.get("alpha", 1.0))
        J0 = np.random.default_rng(xd[1] + ud[0]).uniform(0, 300, P.N)
        Jr, pr = c_oracle.sweep_fused(P, J0)
        for loop in (False, 2):
            J, pi, _ = emu.sweep(P, J0, lanes=lanes, force_generic=loop)
            assert np.array_equal(J, Jr) and np.array_equal(pi, pr), (case, loop)


@pytest.mark.parametrize("order", ["descending", "shuffled"])
def test_emulated_pendulum_with_permuted_action_tables(order):
    """Not ascending B.u: the library (and the emulator's copy of its selection rule) must take the generic loop."""
    case = dict(system="SinglePendulum", x_grid_dim=[33, 41], u_grid_dim=[13], xbar=[-3.14, 0.0], INF=300.0,
                u_lb=[-8.0], u_ub=[8.0], sys_params={"d1": 0.2})
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    perm = np.arange(P.A)[::-1] if order == "descending" else np.random.default_rng(11).permutation(P.A)
    P.tables["bu"][:] = P.tables["bu"][perm]
    P.tables["gu"][:] = P.tables["gu"][perm]
    J0 = np.random.default_rng(5).uniform(0, 300, P.N)
    J, pi, _ = emu.sweep(P, J0, lanes=1)
    Jr, pr = c_oracle.sweep_fused(P, J0)
    assert np.array_equal(J, Jr) and np.array_equal(pi, pr)


@pytest.mark.parametrize("name,case", [
    ("pend_129", dict(system="SinglePendulum", x_grid_dim=[37, 150], u_grid_dim=[41], xbar=[-3.14, 0.0], INF=300.0)),
    ("cartpole_mid", dict(CASES["cartpole_swingup"], x_grid_dim=[7, 9, 13, 17], u_grid_dim=[11])),
    ("twolink_mid", dict(CASES["twolink_soft"], x_grid_dim=[7, 6, 12, 15], u_grid_dim=[5, 7])),
    ("dpend_mid", dict(CASES["dpend_example"], x_grid_dim=[6, 7, 15, 12], u_grid_dim=[7, 5])),
])
def test_emulated_kernels_on_rough_J_equal_c_oracle(name, case):
    """Random J (every corner weight matters), rows longer than a block, a ragged last chunk; lanes = 1 and 4."""
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    J0 = np.random.default_rng(0).uniform(0, 300, P.N)
    Jr, pr = c_oracle.sweep_fused(P, J0)
    for lanes in (1, 4):
        J, pi, _ = emu.sweep(P, J0, lanes=lanes)
        assert np.array_equal(J, Jr) and np.array_equal(pi, pr), lanes


# ---- table-mode kernels (pyro_b200/csrc/table_kernels.cuh) ---------------------------------------------------------
from oracle import np_oracle as npo  # noqa: E402
from tests.cases import POLICY_CASES, oracle_objects  # noqa: E402


@pytest.mark.parametrize("name", ["pend_51x51x11", "cartpole_swingup"])
def test_emulated_lut_kernel_matches_reference_goldens(name):
    """sweep_lut_kernel on reference-identical tables (dynamicprogramming.py:557-570), persistent grid."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    ogrid, ocost = oracle_objects(case)
    x_next, _, _, G = ogrid.tables(ocost)
    P = problem.extract(grid, cf, case.get("alpha", 1.0), force_lut=True)
    J, k = gold["J0"], case["snapshots"][1]
    for _ in range(k):
        J, pi, _ = emu.lut_sweep(P, J, x_next, G)
    assert np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"])


@pytest.mark.parametrize("name", list(POLICY_CASES))
def test_emulated_policy_kernel_matches_reference_goldens(name):
    """sweep_policy_kernel on the reference's own policy-evaluation tables (dynamicprogramming.py:683-752)."""
    case, gold = POLICY_CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0), lut_actions=1)
    J, k = gold["J0"], 0
    for target in case["snapshots"][:2]:
        for _ in range(target - k):
            J_prev = J
            J, pi, st = emu.lut_sweep(P, J_prev, gold["x_next_table"][:, None, :], gold["G"][:, None], grid_blocks=7)
            assert (pi == 0).all() and st[0] == J.max() and st[1] == (J - J_prev).max()
        k = target
        assert np.array_equal(J, gold[f"J_{k}"])


def test_emulated_table_kernels_3d_and_ragged_grids():
    """n = 3 through both table kernels, grids that do not divide the persistent stride."""
    rng = np.random.default_rng(3)
    dims, A = (9, 7, 8), 6
    levels = [np.linspace(-1, 1, dims[0]), np.linspace(0, 3, dims[1]), np.linspace(-2, 5, dims[2])]
    N = int(np.prod(dims))
    X = np.stack([g.reshape(-1) for g in np.meshgrid(*levels, indexing="ij")], axis=1)
    x_next = X[:, None, :] + rng.normal(0, 0.4, (N, A, 3))
    x_next[::17, 0, :] = X[::17]
    x_next[5::19, 1, 2] = 5.0
    G = rng.uniform(0, 1, (N, A))
    J0 = rng.uniform(0, 50, N)

    class Sys3:
        n, m = 3, 1
        x_lb, x_ub = np.array([-1.0, 0.0, -2.0]), np.array([1.0, 3.0, 5.0])
        u_lb, u_ub = np.array([-1.0]), np.array([1.0])

    class Grid3:
        sys, dt = Sys3(), 0.05
        x_grid_dim, u_grid_dim = np.array(dims), np.array([A])
        x_level, u_level = levels, [np.linspace(-1, 1, A)]

    class Cost3:
        INF = 77.0
    J, pi, _ = emu.lut_sweep(problem.extract(Grid3(), Cost3(), 0.97), J0, x_next, G, grid_blocks=5)
    J_ref, pi_ref = npo.lut_sweep(levels, dims, J0, x_next, G, 0.97, use_scipy=True)
    assert np.array_equal(J, J_ref) and np.array_equal(pi, pi_ref)
    J1, pi1, _ = emu.lut_sweep(problem.extract(Grid3(), Cost3(), 0.97, lut_actions=1), J0, x_next[:, 2:3, :], G[:, 2:3], grid_blocks=3)
    J1_ref, _ = npo.lut_sweep(levels, dims, J0, x_next[:, 2:3, :], G[:, 2:3], 0.97, use_scipy=True)
    assert np.array_equal(J1, J1_ref) and (pi1 == 0).all()


@pytest.mark.parametrize("name", ["pend_time_41x61x7", "dpend_example"])
def test_emulated_terminal_cost_kernel(name):
    case = dict(CASES[name], S=[2.0, 0.3] if name.startswith("pend") else [2.0, 0.3, 0.0, 1.5], cost="quadratic")
    case.pop("EPS", None)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    J, pi = emu.terminal(P)
    assert np.array_equal(J, c_oracle.terminal(P)) and (pi == 0).all()


@pytest.mark.parametrize("name", list(POLICY_CASES))
def test_emulated_base_class_policy_evaluator_semantics(name):
    """PolicyEvaluator (the per-node base class, dynamicprogramming.py:636-672) gives exactly INF where the input is
    disallowed, PolicyEvaluatorWithLookUpTable (:700-752) gives INF + alpha*J(x_next) there: the mirror builds the
    tables for either, the policy kernel reproduces both reference classes."""
    from pyro_b200 import dynamicprogramming as dpm
    from tests.cases import LinearFeedback
    case, gold = POLICY_CASES[name], load_golden(name)
    _, grid, cf = build_case(case)

    class NoEngine:
        def __init__(self, N):
            self.N, self.problem = N, type("P", (), {"system_id": 0})()
        def set_J(self, J): pass
        def get_J(self): return np.zeros(self.N)
        def close(self): pass
    kb = case["snapshots"][1]
    for cls, key in ((dpm.PolicyEvaluator, f"Jbase_{kb}"), (dpm.PolicyEvaluatorWithLookUpTable, f"J_{kb}")):
        pe = cls(LinearFeedback(**case["ctl"]), grid, cf, engine_factory=lambda dp, P: NoEngine(grid.nodes_n))
        pe.compute_lookuptable()
        P = problem.extract(grid, cf, case.get("alpha", 1.0), lut_actions=1)
        J = gold["J0"]
        for _ in range(kb):
            J, _, _ = emu.lut_sweep(P, J, pe.x_next_table[:, None, :], pe.G[:, None], grid_blocks=5)
        assert np.array_equal(J, gold[key]), cls.__name__
    assert not np.array_equal(gold[f"Jbase_{kb}"], gold[f"J_{kb}"])   # the two reference classes really differ here


@pytest.mark.parametrize("name,world", [("pend_101x101x21", 4), ("pend_time_41x61x7", 3), ("cartpole_swingup", 3),
                                        ("dpend_example", 2), ("twolink_soft", 3)])
def test_emulated_slab_sweeps_need_only_their_halo(name, world):
    """The multi-GPU data path on the CPU: every rank's launch (the plane ranges pyrodp.cu issues, boundary planes
    first) sees J_next ONLY on its slab + the halo pdp_compute_halo promises — everything else is NaN — and the
    assembled result must equal the whole-grid backup bit for bit, statistics included."""
    import ctypes as C
    from pyro_b200 import _lib, distributed
    case = CASES[name]
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    lo, hi = C.c_int32(), C.c_int32()
    _lib.check(_lib.load().pdp_compute_halo(C.byref(P.c), C.byref(lo), C.byref(hi)))
    lo, hi = lo.value, hi.value
    n0 = P.dims[0]
    plane = P.N // n0
    J0 = np.random.default_rng(1).uniform(0, 200, P.N)
    J_ref, pi_ref, st_ref = emu.sweep(P, J0, lanes=1)
    J = np.full(P.N, -1.0)
    pi = np.full(P.N, -1, dtype=np.int64)
    stats = []
    for r in range(world):
        b, e = distributed.balanced_slab(r, world, n0)
        seen = np.full(P.N, np.nan)                         # what this rank holds: its slab and its halo, nothing else
        a0, a1 = max(0, b - lo), min(n0, e + hi)
        seen[a0 * plane:a1 * plane] = J0[a0 * plane:a1 * plane]
        ranges = [(b, e)] if e - b <= lo + hi else [(b, b + hi), (e - lo, e), (b + hi, e - lo)]   # boundary first, then interior
        for p0, p1 in ranges:
            if p1 > p0:
                stats.append(emu.sweep_planes(P, seen, J, pi, p0, p1, lanes=1))
    assert np.array_equal(J, J_ref) and np.array_equal(pi, pi_ref)
    stats = np.array(stats)
    assert stats[:, 0].max() == st_ref[0] and stats[:, 1].max() == st_ref[1] and stats[:, 2].min() == st_ref[2]


@pytest.mark.parametrize("tile_rows", ["1", "4", "16"])
def test_emulated_range_kernel_block_tiles(monkeypatch, tile_rows):
    """The 4-D range kernel with blocks shaped as tiles of adjacent i2 rows (PYRODP_TILE_ROWS), ragged planes."""
    monkeypatch.setenv("PYRODP_TILE_ROWS", tile_rows)
    for case in (dict(system="CartPole", x_grid_dim=[3, 4, 19, 37], u_grid_dim=[9], xbar=[0.0, float(np.pi), 0.0, 0.0], INF=1000.0),
                 dict(CASES["dpend_example"], x_grid_dim=[3, 4, 21, 35], u_grid_dim=[5, 7])):
        _, grid, cf = build_case(case)
        P = problem.extract(grid, cf, case.get("alpha", 1.0))
        J0 = np.random.default_rng(3).uniform(0, 300, P.N)
        Jr, pr = c_oracle.sweep_fused(P, J0)
        J, pi, st = emu.sweep(P, J0, lanes=1, mech2="range")
        assert np.array_equal(J, Jr) and np.array_equal(pi, pr) and st[0] == Jr.max()


ROLLOUT_FIXTURES = ["pend_51x51x11", "cartpole_swingup", "twolink_soft", "dpend_example"]


def rollout_inputs(name):
    """(Problem, plant parameters, fixture) of a closed-loop rollout fixture (oracle/gen_golden.py: main_rollouts)."""
    case, gold = CASES[name], load_golden("rollout_" + name)
    sys_, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    return P, problem.plant_parameters(sys_, P.system_id), gold


def check_rollout(x, u, gold, rtol=1e-9):
    """Floating-point parity (sin / cos / the 2x2 inverse are evaluated by another library than NumPy's): <= 1e-9 of the
    trajectory's scale at every kept point, for the states and the interpolated inputs."""
    scale = max(1.0, np.abs(gold["x"]).max())
    assert x.shape == gold["x"].shape and u.shape == gold["u"].shape
    assert np.abs(x - gold["x"]).max() <= rtol * scale, np.abs(x - gold["x"]).max()
    assert np.abs(u - gold["u"]).max() <= rtol * max(1.0, np.abs(gold["u"]).max()), np.abs(u - gold["u"]).max()


@pytest.mark.parametrize("name", ROLLOUT_FIXTURES)
def test_emulated_rollout_kernel_matches_reference_closed_loop_trajectories(name):
    """rollout_kernel vs the unmodified reference's `(ctl + sys).compute_trajectory(tf, n, 'euler')`: trajectories that
    leave the grid (controller output 0 there), 2-D value-first and 4-D weight-first policy interpolation."""
    P, phys, gold = rollout_inputs(name)
    npts, tf = int(gold["npts"]), float(gold["tf"])
    x, u = emu.rollout(P, gold["pi"], phys, gold["x0"], npts, tf / (npts - 1))
    check_rollout(x, u, gold)
    xs, us = emu.rollout(P, gold["pi"], phys, gold["x0"], npts, tf / (npts - 1), stride=7)
    assert np.array_equal(xs, x[:, ::7]) and np.array_equal(us, u[:, ::7])


SPLINE_FIXTURES = ["pend_51x51x11", "pend_time_41x61x7"]


def check_spline_snapshot(J, pi, gold, k, rtol=1e-9):
    """Floating-point parity of the spline variant: J within rtol of the reference's scale; pi equal wherever the
    reference's own best and second-best Q differ by more than that tolerance (elsewhere rounding decides the argmin)."""
    J_ref, pi_ref, gap = gold[f"J_{k}"], gold[f"pi_{k}"], gold[f"gap_{k}"]
    scale = np.abs(J_ref).max()
    assert np.abs(J - J_ref).max() <= rtol * scale, (k, np.abs(J - J_ref).max(), scale)
    differ = pi != pi_ref
    assert not (differ & (gap > 10 * rtol * scale)).any(), (k, int(differ.sum()))
    return int(differ.sum())


@pytest.mark.parametrize("name", SPLINE_FIXTURES)
def test_emulated_spline_sweep_matches_reference_spline_class(name):
    """pdp_set_interpolant(SPLINE3): host plan (knots, banded LU), the two fit kernels and sweep_lut_spline_kernel against
    the unmodified reference's DynamicProgramming2DRectBivariateSpline (scipy RectBivariateSpline kx = ky = 3)."""
    from pyro_b200 import dynamicprogramming
    case, gold = CASES[name], load_golden("spline_" + name)
    _, grid, cf = build_case(case, lookup=True)
    alpha = case.get("alpha", 1.0)
    P = problem.extract(grid, cf, alpha, force_lut=True)
    x_next, G = dynamicprogramming.build_lookup_tables(grid, cf, exact_inf=False)
    J, k = gold["J0"], 0
    for target in gold["snapshots"][:2]:
        while k < target:
            J_prev = J
            J, pi, st, coef = emu.spline_sweep(P, J_prev, x_next, G)
            k += 1
            d = J - J_prev
            assert st[0] == J.max() and st[1] == d.max() and st[2] == d.min()
        check_spline_snapshot(J, pi, gold, k)


def test_emulated_spline_fit_equals_scipy_coefficients():
    """The device fit (banded substitutions) against FITPACK's own coefficients on rough data and a non-square grid."""
    from scipy.interpolate import RectBivariateSpline
    from pyro_b200 import dynamicprogramming
    case = dict(system="SinglePendulum", x_grid_dim=[4, 37], u_grid_dim=[3], xbar=[-3.14, 0.0], INF=300.0)
    _, grid, cf = build_case(case, lookup=True)
    P = problem.extract(grid, cf, 0.9, force_lut=True)
    x_next, G = dynamicprogramming.build_lookup_tables(grid, cf, exact_inf=False)
    J0 = np.random.default_rng(3).uniform(-50, 300, P.N)
    J, pi, _, coef = emu.spline_sweep(P, J0, x_next, G)
    sp = RectBivariateSpline(grid.x_level[0], grid.x_level[1], J0.reshape(grid.x_grid_dim), kx=3, ky=3)
    assert np.abs(coef - sp.get_coeffs()).max() <= 1e-10 * np.abs(sp.get_coeffs()).max()
    Q = G + 0.9 * sp(x_next[:, :, 0].flatten(), x_next[:, :, 1].flatten(), grid=False).reshape(G.shape)
    assert np.abs(J - Q.min(axis=1)).max() <= 1e-10 * np.abs(Q.min(axis=1)).max()


class EmuSplineEngine:
    """Engine interface of a table-mode handle in spline mode, backed by the emulated kernels (host-logic test of the
    DynamicProgramming2DRectBivariateSpline mirror without a GPU)."""

    def __init__(self, P):
        self.problem, self.N = P, P.N
        self.J = self.J_next = self.pi = None
        self.launch_count, self.interpolant = 0, "linear"
        self.kernel_info = "sweep_lut_spline_kernel (emulated)"

    def set_lut(self, x_next, G):
        self.x_next, self.G = x_next, G

    def set_interpolant(self, which):
        self.interpolant = which

    def set_J(self, J):
        self.J = np.array(J, dtype=float)

    def get_J(self):
        return self.J.copy()

    def get_J_next(self):
        return self.J_next.copy()

    def get_pi(self):
        return self.pi.copy()

    def sweep(self, n=1):
        assert self.interpolant == "spline3"
        out = np.empty((n, 3))
        for k in range(n):
            self.J_next = self.J
            self.J, self.pi, out[k], _ = emu.spline_sweep(self.problem, self.J_next, self.x_next, self.G)
            self.launch_count += 3
        return out

    def close(self):
        pass


def test_spline_mirror_class_drives_a_table_mode_engine():
    from pyro_b200 import dynamicprogramming
    name = "pend_time_41x61x7"
    case, gold = CASES[name], load_golden("spline_" + name)
    _, grid, cf = build_case(case, lookup=True)
    dp = dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid, cf, engine_factory=lambda dp, P: EmuSplineEngine(P))
    dp.verbose = False
    assert dp._engine.problem.system_id == 0 and dp._engine.interpolant == "spline3"      # forced table mode
    assert np.array_equal(dp.J, gold["J0"])
    k = int(gold["snapshots"][0])
    dp.compute_steps(k)
    check_spline_snapshot(dp.J, dp.pi, gold, k)
    assert dp.k == k and dp.J_next.shape == dp.J.shape
    _, grid4, cf4 = build_case(CASES["twolink_9"])
    with pytest.raises(NotImplementedError):
        dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid4, cf4, engine_factory=lambda dp, P: EmuSplineEngine(P))


# ---- a 3-D example of the reference (obstacles in isavalidstate): table mode, n = 3 -----------------------------------
class EmuLutEngine:
    """Engine interface of a table-mode handle backed by the emulated sweep_lut_kernel (host-logic tests without a GPU)."""
    kernel_info = "sweep_lut_kernel (emulated)"

    def __init__(self, P):
        self.problem, self.N, self.launch_count = P, P.N, 0
        self.J = self.J_next = self.pi = None

    def set_lut(self, x_next, G):
        self.x_next, self.G = np.array(x_next, dtype=float), np.array(G, dtype=float)

    def set_J(self, J):
        self.J = np.array(J, dtype=float)

    def get_J(self):
        return self.J.copy()

    def get_J_next(self):
        return self.J_next.copy()

    def get_pi(self):
        return self.pi.copy()

    def sweep(self, n=1):
        out = np.empty((n, 3))
        for k in range(n):
            self.J_next = self.J
            self.J, self.pi, out[k] = emu.lut_sweep(self.problem, self.J_next, self.x_next, self.G)
            self.launch_count += 1
        return out

    def get_input_from_policy(self, k):
        U = np.stack([g.reshape(-1) for g in np.meshgrid(*[self.problem.tables[f"u_level{i}"] for i in range(self.problem.m)],
                                                         indexing="ij")], axis=1)
        return U[self.pi, k]

    def close(self):
        pass


def test_emulated_lut_kernel_on_the_tables_of_a_3d_reference_example():
    """sweep_lut_kernel<3, G> on the reference's own x_next_table / G of helicopter_tunnel.py (coarse grid): J and pi of the
    reference's DynamicProgrammingWithLookUpTable bit for bit (fixture only; no reference import)."""
    gold = load_golden("helicopter_tunnel_15x13x11")
    dims, udims = [int(d) for d in gold["x_grid_dim"]], [int(d) for d in gold["u_grid_dim"]]
    lb, ub = [-60.0, 0.0, 0.0], [60.0, 10.0, 20.0]

    class Sys3:
        n, m = 3, 1
        x_lb, x_ub, u_lb, u_ub = np.array(lb), np.array(ub), np.array([-20.0]), np.array([20.0])

    class Grid3:
        sys, dt = Sys3(), 0.05
        x_grid_dim, u_grid_dim = np.array(dims), np.array(udims)
        x_level = [np.linspace(lb[i], ub[i], dims[i]) for i in range(3)]
        u_level = [np.linspace(-20.0, 20.0, udims[0])]

    class Cost3:
        INF = 100000
    P = problem.extract(Grid3(), Cost3(), float(gold["alpha"]))
    assert P.system_id == 0
    J, k = gold["J0"], 0
    for target in gold["snapshots"]:
        while k < target:
            J, pi, _ = emu.lut_sweep(P, J, gold["x_next_table"], gold["G"])
            k += 1
        assert np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"]), k


def test_mirror_planner_drops_in_on_the_real_pyro_objects_of_a_3d_example():
    """The reference's helicopter_tunnel.py with only the planner class swapped: real pyro system (obstacles), real
    GridDynamicSystem with its tables, real QuadraticCostFunctionWithDomainCheck -> the mirror selects table mode, builds
    G with the table class's INF semantics and reproduces the reference's J / pi; the controller it hands out is pyro's."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("unmodified reference not present")
    from pyro_b200 import dynamicprogramming
    from tests.cases import helicopter_tunnel_example
    ns = ref_loader.load()
    from pyro.dynamic import drone
    gold = load_golden("helicopter_tunnel_15x13x11")
    with ref_loader.quiet():
        sys_, grid, qcf = helicopter_tunnel_example(drone, ns.costfunction, ns.discretizer)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, qcf, engine_factory=lambda dp, P: EmuLutEngine(P))
    dp.alpha, dp.verbose = float(gold["alpha"]), False
    assert dp._engine.problem.system_id == 0                      # obstacles in isavalidstate: table mode
    assert np.array_equal(dp.J, gold["J0"])
    assert np.array_equal(dp._engine.G, gold["G"]) and np.array_equal(dp._engine.x_next, gold["x_next_table"])
    k = 0
    for target in gold["snapshots"]:
        dp.compute_steps(int(target) - k)
        k = int(target)
        assert np.array_equal(dp.J, gold[f"J_{k}"]) and np.array_equal(dp.pi, gold[f"pi_{k}"]), k
    ctl = dp.get_lookup_table_controller()
    assert type(ctl).__mro__[1].__name__ == "LookUpTableController" and type(ctl).__mro__[1].__module__.startswith("pyro.")
    x = np.array([1.0, 6.0, 5.0])
    with ref_loader.quiet():
        ref = ns.dynamicprogramming.LookUpTableController(grid, gold[f"pi_{k}"])
    assert np.array_equal(ctl.c(x, 0), ref.c(x, 0))


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_table_builder_equals_reference_tables(name):
    """build_tables_kernel (pdp_build_tables) against the reference's own x_next_table / x_next_isok / G samples."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    xn, ok, G = emu.build_tables(P)
    st = int(gold["table_stride"])
    assert np.array_equal(xn[::st], gold["x_next_sample"]) and np.array_equal(ok[::st], gold["x_next_isok_sample"])
    assert np.array_equal(G[::st], gold["G_sample"])
    part = emu.build_tables(P, 7, 5)                                  # a node range
    assert np.array_equal(part[0], xn[7:12]) and np.array_equal(part[2], G[7:12])


def test_emulated_pendulum_kernel_at_full_size_config2_equals_c_oracle_on_every_node():
    """BASELINE config 2 (SinglePendulum 1001 x 1001 x 201, 2.0e8 evals) through the shipped kernel under the emulator: all
    1 002 001 nodes against the C oracle, rough J (about 25 s)."""
    case = dict(system="SinglePendulum", x_grid_dim=[1001, 1001], u_grid_dim=[201], xbar=[-3.14, 0.0], INF=300.0)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    J0 = np.random.default_rng(1).uniform(0, 250, P.N)
    J, pi, st = emu.sweep(P, J0, lanes=1)
    Jr, pr = c_oracle.sweep_fused(P, J0)
    assert np.array_equal(J, Jr) and np.array_equal(pi, pr)
    assert st[0] == Jr.max() and st[1] == (Jr - J0).max() and st[2] == (Jr - J0).min()
