"""Host-side logic and the C-ABI surface, no GPU compute."""
import ctypes
import io
import os
import re
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import c_oracle, ref_loader
from pyro_b200 import _lib, costfunction, discretizer, distributed, dynamicprogramming, problem, systems
from tests.cases import CASES, build_case
from tests.conftest import ROOT, has_gpu, load_golden
from tests.fake_engine import FakeEngine


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pyrodp.h")).read()
    declared = set(re.findall(r"\b(pdp_[a-z_A-Z0-9]+)\s*\(", header))
    declared -= {"pdp_problem", "pdp_handle", "pdp_stats"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pyrodp.h but not exported"
    assert set(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert _lib.load().pdp_abi_version() == _lib.PDP_ABI_VERSION


def test_struct_layout_matches_header():
    """sizeof(pdp_problem) seen by gcc == ctypes.sizeof (guards against field drift)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "pyrodp.h"\nint main(){printf("%zu %zu", sizeof(pdp_problem), sizeof(pdp_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        a, b = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    assert int(a) == ctypes.sizeof(_lib.pdp_problem) and int(b) == ctypes.sizeof(_lib.pdp_stats)


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_create_fails_loudly():
    _, grid, cf = build_case(CASES["pend_51x51x11"])
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        dynamicprogramming.DynamicProgramming(grid, cf)


def test_extract_classification_and_errors():
    _, grid, cf = build_case(CASES["pend_51x51x11"])
    P = problem.extract(grid, cf, 1.0)
    assert P.system_id == _lib.PDP_SYS_PENDULUM and P.cost_id == _lib.PDP_COST_QUADRATIC
    assert (P.N, P.A) == (2601, 11) and P.tables["gu"].shape == (11,)
    assert problem.extract(*build_case(CASES["cartpole_swingup"])[1:], 1.0).system_id == _lib.PDP_SYS_CARTPOLE
    assert problem.extract(*build_case(CASES["twolink_9"])[1:], 1.0).system_id == _lib.PDP_SYS_TWOLINK
    assert problem.extract(*build_case(CASES["pend_time_41x61x7"])[1:], 0.9).cost_id == _lib.PDP_COST_TIME

    class Obstacle(systems.SinglePendulum):  # custom validity -> generic LUT mode
        def isavalidstate(self, x):
            return super().isavalidstate(x) and abs(x[0]) > 0.1
    g2 = discretizer.GridDynamicSystem(Obstacle(), [11, 11], [3])
    assert problem.extract(g2, cf, 1.0).system_id == _lib.PDP_SYS_LUT

    class Custom(costfunction.CostFunction):
        xbar = np.zeros(2)
    assert problem.extract(grid, Custom(), 1.0).system_id == _lib.PDP_SYS_LUT
    with pytest.raises(NotImplementedError):
        problem.extract(grid, cf, 1.0, interpol_method="cubic")
    cf_bad = costfunction.QuadraticCostFunction(3, 1)
    with pytest.raises(ValueError):
        problem.extract(grid, cf_bad, 1.0)
    s5 = systems.MechanicalSystem(3)
    with pytest.raises(NotImplementedError):
        discretizer.GridDynamicSystem(s5, [3] * 6, [3] * 3)


def test_fingerprint_tracks_parameters_not_pointers():
    _, grid, cf = build_case(CASES["pend_51x51x11"])
    a, b = problem.extract(grid, cf, 1.0), problem.extract(grid, cf, 1.0)
    assert a.fingerprint() == b.fingerprint()
    cf.INF = 301.0
    assert problem.extract(grid, cf, 1.0).fingerprint() != a.fingerprint()
    cf.INF = 300.0
    assert problem.extract(grid, cf, 0.99).fingerprint() != a.fingerprint()


@pytest.mark.skipif(not ref_loader.available(), reason="reference not present (GPU box)")
@pytest.mark.parametrize("name", ["pend_51x51x11", "dpend_example", "twolink_soft", "cartpole_swingup", "pend_time_41x61x7"])
def test_extract_from_real_pyro_objects_equals_mirrors(name):
    """Drop-in: the descriptor read from unmodified pyro objects == the one read from the mirrors."""
    import importlib
    gen = importlib.import_module("oracle.gen_golden")
    ns = ref_loader.load()
    case = dict(CASES[name], x_grid_dim=[5] * len(CASES[name]["x_grid_dim"]))  # tiny grid: table build is a Python loop
    with ref_loader.quiet():
        _, rgrid, rcf, rdp = gen.build_reference(ns, case)
    _, grid, cf = build_case(case)
    A = problem.extract(rgrid, rcf, rdp.alpha)
    B = problem.extract(grid, cf, case.get("alpha", 1.0))
    assert A.fingerprint() == B.fingerprint()
    assert A.system_id == B.system_id != _lib.PDP_SYS_LUT
    for k in A.tables:
        assert np.array_equal(A.tables[k], B.tables[k]), k


def test_grid_mirror_matches_reference_layout():
    case = CASES["dpend_example"]
    _, grid, _ = build_case(case)
    gold = load_golden("dpend_example")
    assert grid.nodes_n == 11 * 9 * 13 * 11 and grid.actions_n == 15
    # C order, last axis fastest (discretizer.py:223-241)
    s = grid.state_from_node_id
    assert s[1, 3] == grid.x_level[3][1] and s[11, 2] == grid.x_level[2][1]
    assert np.array_equal(grid.index_from_node_id[11 * 13 + 5], [0, 1, 0, 5])
    assert grid.node_id_from_index[1, 2, 3, 4] == ((1 * 9 + 2) * 13 + 3) * 11 + 4
    assert np.array_equal(grid.input_from_action_id[7], [grid.u_level[0][1], grid.u_level[1][2]])
    pi = gold["pi_30"]
    want = np.array([grid.input_from_action_id[a, 1] for a in pi])
    assert np.array_equal(grid.get_input_from_policy(pi, 1), want)
    with pytest.raises(ValueError):
        grid.compute_interpolation_function(np.zeros(7))
    assert grid.get_nearest_action_id_from_input(np.array([0.0, 0.0])) == 1 * 5 + 2


def fake_factory(dp, P):
    return FakeEngine(P)


def test_dp_bookkeeping_matches_reference_goldens_with_standin_engine():
    """compute_steps / history / printed line / J,pi attributes, engine replaced by the CPU stand-in."""
    case, gold = CASES["pend_51x51x11"], load_golden("pend_51x51x11")
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgramming(grid, cf, engine_factory=fake_factory)
    assert dp.save_time_history and len(dp.J_list) == 1 and dp.k == 0 and dp.t == 0
    assert dp.pi.dtype == np.int64 and dp.J.dtype == np.float64 and np.array_equal(dp.J, gold["J0"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        dp.compute_steps(2)
    lines = buf.getvalue().strip().splitlines()
    assert lines[0] == "Computing 2 backward DP iterations:"
    assert re.fullmatch(r"1 t:-0\.05 Elasped:\d+\.\d\d max: \d+\.\d\d dmax:\d+\.\d\d dmin:-?\d+\.\d\d", lines[2]), lines[2]
    assert dp.k == 2 and abs(dp.t + 0.1) < 1e-12 and len(dp.J_list) == 3 and len(dp.t_list) == 3
    assert np.array_equal(dp.J, gold["J_2"]) and np.array_equal(dp.pi, gold["pi_2"])
    assert np.array_equal(dp.J_next, gold["J_1"]) and np.array_equal(dp.J_list[1], gold["J_1"])
    # batched path (history off) gives the same numbers
    dp2 = dynamicprogramming.DynamicProgramming(grid, cf, engine_factory=fake_factory)
    dp2.save_time_history, dp2.verbose = False, False
    dp2.compute_steps(10)
    assert dp2.k == 10 and np.array_equal(dp2.J, gold["J_10"]) and np.array_equal(dp2.pi, gold["pi_10"])


def test_solve_bellman_equation_stops_like_reference():
    case = CASES["pend_time_41x61x7"]
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf, engine_factory=fake_factory)
    dp.alpha, dp.verbose = 0.9, False
    dp.solve_bellman_equation(tol=0.5)
    # the reference loop: sweep until max(|dmax|,|dmin|) <= tol (dynamicprogramming.py:303-308)
    P = problem.extract(grid, cf, 0.9)
    J, k = c_oracle.terminal(P), 0
    while True:
        Jn = J
        J, pi = c_oracle.sweep_fused(P, Jn)
        k += 1
        if max(abs((J - Jn).max()), abs((J - Jn).min())) <= 0.5:
            break
    assert dp.k == k and np.array_equal(dp.J, J) and np.array_equal(dp.pi, pi)
    # alpha change after construction is picked up lazily and J is carried over
    dp.alpha = 0.5
    dp.compute_steps(1)
    J2, _ = c_oracle.sweep_fused(problem.extract(grid, cf, 0.5), J)
    assert np.array_equal(dp.J, J2)
    # guard for the non-converging alpha=1 case (hard part 7)
    dp3 = dynamicprogramming.DynamicProgramming(*build_case(CASES["pend_51x51x11"])[1:], engine_factory=fake_factory)
    dp3.verbose, dp3.max_sweeps = False, 3
    dp3.solve_bellman_equation(tol=1e-9)
    assert dp3.k == 3


def test_policy_tools_and_checkpoint_format(tmp_path):
    case, gold = CASES["cartpole_swingup"], load_golden("cartpole_swingup")
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgramming(grid, cf, engine_factory=fake_factory)
    dp.verbose = False
    dp.compute_steps(10)
    ctl = dp.get_lookup_table_controller()
    x = np.array([0.3, 2.0, -0.4, 0.7])
    want = grid.compute_interpolation_function(grid.get_input_from_policy(gold["pi_10"], 0), "linear", False, 0)(x)[0]
    assert ctl.c(x, ctl.rbar)[0] == want
    dp.save_latest(str(tmp_path / "ck"))
    assert np.array_equal(np.load(tmp_path / "ck_J_inf.npy"), dp.J_next)      # saves J_next (dynamicprogramming.py:484)
    assert np.load(tmp_path / "ck_pi_inf.npy").dtype == np.int64
    dp.clean_infeasible_set(tol=1)
    bad = gold["J_10"] > cf.INF - 1
    assert (dp.J[bad] == cf.INF).all() and (dp.pi[bad] == grid.get_nearest_action_id_from_input(grid.sys.ubar)).all()
    assert np.array_equal(dp.J[~bad], gold["J_10"][~bad])


def test_slab_partition():
    for n0, w in [(201, 8), (151, 8), (1001, 4), (5, 8), (16, 2), (7, 1)]:
        slabs = [distributed.slab_of(r, w, n0) for r in range(w)]
        per = slabs[0][2]
        assert per * w >= n0 and all(s[2] == per for s in slabs)
        covered = [p for b, e, _ in slabs for p in range(b, e)]
        assert covered == list(range(n0))
        assert all(b == min(r * per, n0) for r, (b, e, _) in enumerate(slabs))


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under pyro_b200/ (Python or CUDA) may import, include or execute it."""
    import re
    pkg = os.path.join(ROOT, "pyro_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert not re.search(r"#include\s+[\"<][^\">]*oracle", src), f
                assert "dp_oracle" not in src and "np_oracle" not in src and "c_oracle" not in src, f


@pytest.mark.skipif(not ref_loader.available(), reason="reference not present (GPU box)")
@pytest.mark.parametrize("name", ["pend_time_41x61x7", "cartpole_swingup"])
def test_grid_mirror_host_methods_equal_the_real_reference(name, tmp_path):
    """The callers either side of the sweep (discretizer.py:314-664): dense tables, their .npz format, the
    state/input <-> index conversions and the 2-D slice, on the mirror vs the unmodified reference, bit for bit."""
    import importlib
    gen = importlib.import_module("oracle.gen_golden")
    ns = ref_loader.load()
    dims = [6, 5] if len(CASES[name]["x_grid_dim"]) == 2 else [4, 5, 3, 4]
    case = dict(CASES[name], x_grid_dim=dims)
    with ref_loader.quiet():
        rsys, rgrid, _, _ = gen.build_reference(ns, case)
        rgrid.compute_nearest_snext_table()
    _, grid, _ = build_case(case, lookup=True)
    grid.compute_nearest_snext_table()
    for attr in ("x_next_table", "x_next_isok", "action_isok", "s_next_table", "state_from_node_id", "index_from_node_id",
                 "node_id_from_index", "input_from_action_id", "index_from_action_id", "action_id_from_index"):
        assert np.array_equal(getattr(grid, attr), getattr(rgrid, attr)), attr
    rng = np.random.default_rng(4)
    for _ in range(50):
        x = rng.uniform(rsys.x_lb - 1.0, rsys.x_ub + 1.0)
        u = rng.uniform(rsys.u_lb - 1.0, rsys.u_ub + 1.0)
        assert np.array_equal(grid.get_index_from_state(x), rgrid.get_index_from_state(x))
        assert np.array_equal(grid.get_nearest_index_from_state(x), rgrid.get_nearest_index_from_state(x))
        assert grid.get_nearest_node_id_from_state(x) == rgrid.get_nearest_node_id_from_state(x)
        assert np.array_equal(grid.get_index_from_input(u), rgrid.get_index_from_input(u))
        assert np.array_equal(grid.get_nearest_index_from_input(u), rgrid.get_nearest_index_from_input(u))
        assert grid.get_nearest_action_id_from_input(u) == rgrid.get_nearest_action_id_from_input(u)
    Z = rng.uniform(0, 1, grid.nodes_n)
    for a1, a2 in ((0, 1),) if len(dims) == 2 else ((0, 1), (1, 3), (2, 0), (3, 2)):
        want = rgrid.get_2D_slice_of_grid(rgrid.get_grid_from_array(Z), a1, a2)
        assert np.array_equal(grid.get_2D_slice_of_grid(grid.get_grid_from_array(Z), a1, a2), want), (a1, a2)
    # on-disk format: written by one side, read by the other
    grid.save_lookup_tables(str(tmp_path / "mine"))
    rgrid.x_next_table = None
    rgrid.load_lookup_tables(str(tmp_path / "mine"))
    assert np.array_equal(rgrid.x_next_table, grid.x_next_table) and np.array_equal(rgrid.action_isok, grid.action_isok)
    rgrid.save_lookup_tables(str(tmp_path / "theirs"))
    _, fresh, _ = build_case(case)
    fresh.load_lookup_tables(str(tmp_path / "theirs"))
    assert np.array_equal(fresh.x_next_table, grid.x_next_table) and np.array_equal(fresh.x_next_isok, grid.x_next_isok)
    if len(dims) == 2:
        J = rng.uniform(0, 1, grid.nodes_n)
        p = np.array([[0.3, -0.2], [1.1, 0.7]])
        assert np.array_equal(grid.compute_bivariatespline_2D_interpolation_function(J).ev(p[:, 0], p[:, 1]),
                              rgrid.compute_bivariatespline_2D_interpolation_function(J).ev(p[:, 0], p[:, 1]))


@pytest.mark.skipif(not ref_loader.available(), reason="reference not present (GPU box)")
def test_controller_and_infeasible_cleaning_equal_the_real_reference():
    """The step after the sweep (dynamicprogramming.py:27-107, 322-334, 472-477) on the mirror vs the unmodified
    reference: same policy in -> same control law out, same J / pi after clean_infeasible_set."""
    import importlib
    gen = importlib.import_module("oracle.gen_golden")
    ns = ref_loader.load()
    case = dict(CASES["cartpole_swingup"], x_grid_dim=[5, 7, 5, 7])
    with ref_loader.quiet():
        rsys, rgrid, rcf, rdp = gen.build_reference(ns, case)
        rdp.compute_steps(6)
        rctl = rdp.get_lookup_table_controller()
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf, engine_factory=fake_factory)
    dp.verbose = False
    dp.compute_steps(6)
    assert np.array_equal(dp.J, rdp.J) and np.array_equal(dp.pi, rdp.pi)
    ctl = dp.get_lookup_table_controller()
    rng = np.random.default_rng(8)
    for _ in range(40):
        x = rng.uniform(rsys.x_lb - 0.5, rsys.x_ub + 0.5)     # some states outside the grid: fill value 0
        assert np.array_equal(ctl.c(x, ctl.rbar), rctl.c(x, rctl.rbar))
        assert np.array_equal(ctl.cbar(x), rctl.cbar(x))
    with ref_loader.quiet():
        rdp.clean_infeasible_set(tol=1)
    dp.clean_infeasible_set(tol=1)
    assert np.array_equal(dp.J, rdp.J) and np.array_equal(dp.pi, rdp.pi)


def test_classify_routes_overriding_subclasses_to_lut_mode():
    """a subclass that overrides the dynamics or the cost must NOT inherit its parent's fused kernel."""
    from pyro_b200 import _lib, costfunction, systems

    class G:
        pass

    def ids(sys_, cf=None):
        g = G()
        g.sys = sys_
        return problem.classify(g, cf or costfunction.QuadraticCostFunction.from_sys(sys_))

    class Renamed(systems.SinglePendulum):          # parameters only: still the parent's equations
        def __init__(self):
            super().__init__()
            self.m1 = 2.0

    class FlippedGravity(systems.SinglePendulum):   # the reference's InvertedPendulum does exactly this (pendulum.py:283)
        def g(self, q):
            return -systems.SinglePendulum.g(self, q)

    class OtherB(systems.DoublePendulum):           # ... and its Acrobot this (pendulum.py:699)
        def B(self, q):
            return np.array([[0.0], [1.0]])

    class Obstacle(systems.CartPole):
        def isavalidstate(self, x):
            return bool(super().isavalidstate(x) and abs(x[0]) > 0.1)

    class MyCost(costfunction.QuadraticCostFunction):
        def g(self, x, u, t=0):
            return 2.0 * super().g(x, u, t)

    assert ids(systems.SinglePendulum()) == (_lib.PDP_SYS_PENDULUM, _lib.PDP_COST_QUADRATIC)
    assert ids(Renamed()) == (_lib.PDP_SYS_PENDULUM, _lib.PDP_COST_QUADRATIC)
    for s in (FlippedGravity(), OtherB(), Obstacle()):
        assert ids(s) == (_lib.PDP_SYS_LUT, 0), type(s).__name__
    assert ids(systems.SinglePendulum(), MyCost(2, 1)) == (_lib.PDP_SYS_LUT, 0)
    patched = systems.CartPole()
    patched.f = lambda x, u, t=0: np.zeros(4)       # a method patched onto the instance
    assert ids(patched) == (_lib.PDP_SYS_LUT, 0)
    if ref_loader.available():
        ns = ref_loader.load()
        real = lambda s: ids(s, ns.costfunction.QuadraticCostFunction.from_sys(s))
        assert real(ns.pendulum.SinglePendulum())[0] == _lib.PDP_SYS_PENDULUM
        assert real(ns.pendulum.DoublePendulum())[0] == _lib.PDP_SYS_TWOLINK
        assert real(ns.manipulator.TwoLinkManipulator())[0] == _lib.PDP_SYS_TWOLINK
        assert real(ns.cartpole.CartPole())[0] == _lib.PDP_SYS_CARTPOLE
        assert real(ns.pendulum.InvertedPendulum())[0] == _lib.PDP_SYS_LUT
        assert real(ns.pendulum.Acrobot())[0] == _lib.PDP_SYS_LUT


def test_lookup_table_controller_has_the_static_controller_surface():
    """what reference scripts do with dp.get_lookup_table_controller() (controller.py:22-162)."""
    case = dict(CASES["pend_51x51x11"], x_grid_dim=[9, 7], u_grid_dim=[3])
    _, grid, cf = build_case(case)
    pi = np.arange(grid.nodes_n) % 3
    ctl = dynamicprogramming.LookUpTableController(grid, pi)
    assert (ctl.k, ctl.m, ctl.p) == (1, 1, 2) and ctl.name == "Tabular Controller"
    assert ctl.ref_label == ["Ref. 0"] and ctl.ref_units == [""] and ctl.r_ub[0] == 10 and ctl.r_lb[0] == -10
    x = np.array([0.3, -0.7])
    assert np.array_equal(ctl.cbar(x), ctl.c(x, ctl.t2r(0.0))) and ctl.forward_kinematic_lines_plus(x, 0, 0) == (None, None, None)
    if ref_loader.available():
        ns = ref_loader.load()
        cl = ctl + ns.pendulum.SinglePendulum()                         # pyro's own ClosedLoopSystem
        assert type(cl).__name__ == "ClosedLoopSystem" and cl.controller is ctl
        # with pyro importable the planner hands out pyro's LookUpTableController itself, fed with the device tables
        rc = dynamicprogramming.make_reference_controller(grid, pi, [grid.get_input_from_policy(pi, 0)])
        assert isinstance(rc, ns.dynamicprogramming.LookUpTableController) and np.array_equal(rc.c(x, rc.rbar), ctl.c(x, ctl.rbar))


def test_table_fallback_reproduces_both_inf_semantics():
    """base class = exact INF on a disallowed input (dynamicprogramming.py:230-233), table class =
    INF + alpha*J(x_next) (:545-549,:567); f and g are evaluated at the t that is passed."""
    from pyro_b200 import systems, costfunction, discretizer

    class Picky(systems.SinglePendulum):
        def isavalidinput(self, x, u):
            return bool(super().isavalidinput(x, u) and not (x[0] > 0 and u[0] > 0))

        def f(self, x, u, t=0):
            return super().f(x, u, t) * (1.0 + t)

    sys_ = Picky()
    grid = discretizer.GridDynamicSystem(sys_, [7, 5], [3], 0.05)
    cf = costfunction.QuadraticCostFunction.from_sys(sys_)
    xa, Ga = dynamicprogramming.build_lookup_tables(grid, cf, t=0.0, exact_inf=True)
    xb, Gb = dynamicprogramming.build_lookup_tables(grid, cf, t=0.0, exact_inf=False)
    assert np.array_equal(Ga, Gb)
    X, U = grid.state_from_node_id, grid.input_from_action_id
    bad = np.array([[not sys_.isavalidinput(X[s], U[a]) for a in range(3)] for s in range(grid.nodes_n)])
    assert bad.any() and (Ga[bad] == cf.INF).all()
    assert (xa[bad] > sys_.x_ub).all() and np.array_equal(xa[~bad], xb[~bad])     # moved outside the box: RGI returns 0 there
    assert not (xb[bad] > sys_.x_ub).all()
    xt, _ = dynamicprogramming.build_lookup_tables(grid, cf, t=1.0, exact_inf=False)
    s, a = 3, 1
    assert np.array_equal(xt[s, a], sys_.f(X[s], U[a], 1.0) * grid.dt + X[s]) and not np.array_equal(xt[s, a], xb[s, a])
