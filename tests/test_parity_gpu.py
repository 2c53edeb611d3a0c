"""GPU parity: the CUDA path, called through the C ABI, against the reference goldens and the oracle.

Bars (BASELINE.json north_star): J within 1e-5 relative (L-inf over max|J_ref|), pi index-exact.
The kernels restate the reference's IEEE operations one by one, so the tests additionally record
whether J is bit-identical (it is expected to be when the box's NumPy produces the same table
bits as the container that generated the goldens).
"""
import numpy as np
import pytest

from oracle import c_oracle, np_oracle as npo
from pyro_b200 import _lib, dynamicprogramming, problem
from pyro_b200.engine import Engine
from tests.cases import CASES, POLICY_CASES, LinearFeedback, build_case, oracle_objects
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # north_star tolerance on J


def rel_err(J, J_ref):
    scale = np.abs(J_ref).max()
    return np.abs(J - J_ref).max() / (scale if scale > 0 else 1.0)


@pytest.mark.parametrize("name", list(CASES))
def test_fused_kernels_match_reference_goldens(name):
    """Public API (DynamicProgramming.compute_steps) vs tier-0 fixtures at every snapshot."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
    dp.alpha, dp.verbose = case.get("alpha", 1.0), False
    assert np.array_equal(dp.J, gold["J0"])
    k = 0
    for target in case["snapshots"]:
        dp.compute_steps(target - k)
        k = target
        J_ref, pi_ref = gold[f"J_{k}"], gold[f"pi_{k}"]
        assert dp.J.dtype == np.float64 and dp.pi.dtype == np.int64 and dp.J.shape == (grid.nodes_n,)
        err, mism = rel_err(dp.J, J_ref), int((dp.pi != pi_ref).sum())
        print(f"{name} k={k}: rel err {err:.2e}, pi mismatches {mism}, bit-exact {np.array_equal(dp.J, J_ref)}")
        assert err <= RTOL and mism == 0
    assert dp._engine.launch_count >= case["snapshots"][-1]


@pytest.mark.parametrize("name", ["pend_51x51x11", "dpend_example", "cartpole_swingup"])
def test_lut_mode_kernel_is_bit_exact(name):
    """LUT-mode kernel (generic fallback, dynamicprogramming.py:557-570) on reference-identical tables: exact."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    ogrid, ocost = oracle_objects(case)
    x_next, _, _, G = ogrid.tables(ocost)  # bit-identical to the reference's tables (tests/test_oracle.py)
    eng = Engine(problem.extract(grid, cf, case.get("alpha", 1.0), force_lut=True))
    eng.set_lut(x_next, G)
    eng.set_J(gold["J0"])
    k = 0
    for target in case["snapshots"][:3]:
        stats = eng.sweep(target - k)
        k = target
        assert np.array_equal(eng.get_J(), gold[f"J_{k}"]) and np.array_equal(eng.get_pi(), gold[f"pi_{k}"])
    d = gold[f"J_{k}"] - eng.get_J_next()
    assert stats[-1, 0] == gold[f"J_{k}"].max() and stats[-1, 1] == d.max() and stats[-1, 2] == d.min()
    eng.close()


def test_lut_mode_3d_generic_system():
    """n = 3 has no fused kernel: a synthetic 3-state system through LUT mode vs the NumPy RGI restatement."""
    rng = np.random.default_rng(3)
    dims, A = (9, 7, 8), 6
    levels = [np.linspace(-1, 1, dims[0]), np.linspace(0, 3, dims[1]), np.linspace(-2, 5, dims[2])]
    N = int(np.prod(dims))
    X = np.stack([g.reshape(-1) for g in np.meshgrid(*levels, indexing="ij")], axis=1)
    x_next = X[:, None, :] + rng.normal(0, 0.4, (N, A, 3))
    x_next[::17, 0, :] = X[::17]  # exact node hits
    x_next[5::19, 1, 2] = 5.0     # exactly on the upper bound
    oob = np.any((x_next < [-1, 0, -2]) | (x_next > [1, 3, 5]), axis=-1)
    G = np.where(oob, 77.0, rng.uniform(0, 1, (N, A)))
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
    eng = Engine(problem.extract(Grid3(), Cost3(), 0.97))
    assert eng.problem.system_id == _lib.PDP_SYS_LUT
    eng.set_lut(x_next, G)
    eng.set_J(J0)
    eng.sweep(1)
    J_ref, pi_ref = npo.lut_sweep(levels, dims, J0, x_next, G, 0.97, use_scipy=True)
    assert np.array_equal(eng.get_J(), J_ref) and np.array_equal(eng.get_pi(), pi_ref)
    eng.close()
    # one column per node (policy evaluation, dynamicprogramming.py:743-752): the streaming kernel, n = 3
    pe = Engine(problem.extract(Grid3(), Cost3(), 0.97, lut_actions=1))
    pe.set_lut(x_next[:, 2:3, :], G[:, 2:3])
    pe.set_J(J0)
    st = pe.sweep(2)
    J1, _ = npo.lut_sweep(levels, dims, J0, x_next[:, 2:3, :], G[:, 2:3], 0.97, use_scipy=True)
    J2, _ = npo.lut_sweep(levels, dims, J1, x_next[:, 2:3, :], G[:, 2:3], 0.97, use_scipy=True)
    assert np.array_equal(pe.get_J(), J2) and (pe.get_pi() == 0).all() and st[1, 0] == J2.max() and st[1, 1] == (J2 - J1).max()
    pe.close()


@pytest.mark.parametrize("n", [2, 4])
def test_policy_kernel_on_a_large_grid_equals_numpy_restatement(n):
    """The streaming policy-evaluation kernel at a size with many trips per thread and a ragged tail."""
    rng = np.random.default_rng(9)
    dims = (307, 293) if n == 2 else (23, 19, 21, 17)
    levels = [np.linspace(-1.0 - d, 2.0 + d, dims[d]) for d in range(n)]
    N = int(np.prod(dims))
    X = np.stack([g.reshape(-1) for g in np.meshgrid(*levels, indexing="ij")], axis=1)
    step = np.array([levels[d][1] - levels[d][0] for d in range(n)])
    x_next = X[:, None, :] + rng.uniform(-2.2, 2.2, (N, 1, n)) * step
    G = rng.uniform(0, 2, (N, 1))

    class SysN:
        m = 1
        x_lb, x_ub = np.array([l[0] for l in levels]), np.array([l[-1] for l in levels])
        u_lb, u_ub = np.array([-1.0]), np.array([1.0])
    SysN.n = n

    class GridN:
        sys, dt = SysN(), 0.05
        x_grid_dim, u_grid_dim = np.array(dims), np.array([3])
        x_level, u_level = levels, [np.linspace(-1, 1, 3)]

    class CostN:
        INF = 50.0
    eng = Engine(problem.extract(GridN(), CostN(), 0.9, lut_actions=1))
    eng.set_lut(x_next, G)
    J0 = rng.uniform(0, 100, N)
    eng.set_J(J0)
    eng.sweep(1)
    J_ref, _ = npo.lut_sweep(levels, list(dims), J0, x_next, G, 0.9, use_scipy=False)
    assert np.array_equal(eng.get_J(), J_ref)
    eng.close()


MID = {
    "pend_301": dict(system="SinglePendulum", x_grid_dim=[301, 257], u_grid_dim=[41], xbar=[-3.14, 0.0], INF=300.0),
    "dpend_21": dict(CASES["dpend_example"], x_grid_dim=[21, 19, 23, 21], u_grid_dim=[7, 5]),
    "twolink_21": dict(CASES["twolink_soft"], x_grid_dim=[21, 17, 21, 25], u_grid_dim=[5, 7]),
    "cartpole_25": dict(CASES["cartpole_swingup"], x_grid_dim=[15, 25, 21, 27], u_grid_dim=[11]),
}


@pytest.mark.parametrize("name", list(MID))
def test_mid_size_random_J_equals_c_oracle(name, monkeypatch, generic=0):
    """Sizes the reference cannot build in reasonable time; rough (random) J so every corner weight matters."""
    monkeypatch.setenv("PYRODP_GENERIC", str(generic))
    case = MID[name]
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    eng = Engine(P)
    J0 = np.random.default_rng(0).uniform(0, 300, P.N)
    eng.set_J(J0)
    stats = eng.sweep(2)
    J1, pi1 = c_oracle.sweep_fused(P, J0)
    J2, pi2 = c_oracle.sweep_fused(P, J1)
    J, pi = eng.get_J(), eng.get_pi()
    print(f"{name}: rel err {rel_err(J, J2):.2e}, pi mismatches {(pi != pi2).sum()}, bit-exact {np.array_equal(J, J2)}")
    assert np.array_equal(J, J2) and np.array_equal(pi, pi2)
    assert np.array_equal(eng.get_J_next(), J1)
    d = J2 - J1
    assert stats[1, 0] == J2.max() and stats[1, 1] == d.max() and stats[1, 2] == d.min()
    eng.close()


def test_pendulum_order_agnostic_loop_on_mid_size_and_descending_inputs(monkeypatch):
    """The generic pendulum loop (any action order): forced on an ascending grid, and selected by the library
    itself when B.u is not ascending (here: the descriptor's per-action tables reversed and shuffled)."""
    test_mid_size_random_J_equals_c_oracle("pend_301", monkeypatch, generic=1)
    monkeypatch.setenv("PYRODP_GENERIC", "0")
    case = dict(system="SinglePendulum", x_grid_dim=[65, 77], u_grid_dim=[23], xbar=[-3.14, 0.0], INF=300.0,
                u_lb=[-8.0], u_ub=[8.0], sys_params={"d1": 0.2})
    _, grid, cf = build_case(case)
    for order in ("descending", "shuffled"):
        P = problem.extract(grid, cf, 1.0)
        perm = np.arange(P.A)[::-1] if order == "descending" else np.random.default_rng(11).permutation(P.A)
        # the per-action tables of the descriptor (B.u, du'R du) in another order: engine and oracle read the same bytes
        P.tables["bu"][:] = P.tables["bu"][perm]
        P.tables["gu"][:] = P.tables["gu"][perm]
        eng = Engine(P)
        J0 = np.random.default_rng(5).uniform(0, 300, P.N)
        eng.set_J(J0)
        eng.sweep(1)
        J1, pi1 = c_oracle.sweep_fused(P, J0)
        assert np.array_equal(eng.get_J(), J1) and np.array_equal(eng.get_pi(), pi1), order
        eng.close()


@pytest.mark.parametrize("loop", ["1", "2"])
def test_pendulum_loop_nest_and_pair_loop_on_edge_shapes(loop, monkeypatch):
    """Both MONO loops of sweep_pendulum_kernel (PYRODP_PEND_LOOP: 1 = pair loop, shipped; 2 = loop nest) where
    their special cases live — padding records, cells narrower than an action step, parked lanes, damping, two-level
    axes — on rough J against the C oracle, for every lane split the library may choose."""
    monkeypatch.setenv("PYRODP_PEND_LOOP", loop)
    shapes = [([19, 260], [1]), ([19, 260], [2]), ([23, 300], [3]), ([17, 131], [37]), ([9, 40], [201]), ([5, 3], [7]), ([5, 2], [5]),
              ([301, 517], [64])]
    extras = [dict(), dict(sys_params={"d1": 0.3}, x_lb=[-2.0, -1.5], x_ub=[1.0, 2.5]), dict(alpha=0.9, x_lb=[-3.0, -0.4], x_ub=[3.0, 0.4])]
    for (xd, ud) in shapes:
        for extra in extras:
            case = dict(system="SinglePendulum", x_grid_dim=xd, u_grid_dim=ud, xbar=[-3.14, 0.0], INF=300.0, **extra)
            _, grid, cf = build_case(case)
            P = problem.extract(grid, cf, case.get("alpha", 1.0))
            J0 = np.random.default_rng(xd[1] + ud[0]).uniform(0, 300, P.N)
            Jr, pr = c_oracle.sweep_fused(P, J0)
            for lanes in (0, 1, 4, 16):
                if lanes > 1 and ud[0] < 4 * lanes:
                    continue
                monkeypatch.setenv("PYRODP_LANES", str(lanes))
                eng = Engine(P)
                want = "loop nest" if loop == "2" else "pair loop"
                assert want in eng.kernel_info, eng.kernel_info
                eng.set_J(J0)
                eng.sweep(1)
                assert np.array_equal(eng.get_J(), Jr) and np.array_equal(eng.get_pi(), pr), (case, lanes)
                eng.close()


def test_full_size_config2_sampled_against_oracle_and_properties():
    """BASELINE config 2 (SinglePendulum 1001x1001x201) at full size: all 1 002 001 nodes against the C
    oracle, plus size-independent properties (determinism, monotonicity of the Bellman operator,
    constant-shift equivariance for alpha = 1 on nodes whose minimiser is a valid transition)."""
    case = dict(system="SinglePendulum", x_grid_dim=[1001, 1001], u_grid_dim=[201], xbar=[-3.14, 0.0], INF=300.0)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    eng = Engine(P)
    rng = np.random.default_rng(1)
    J0 = rng.uniform(0, 250, P.N)
    eng.set_J(J0)
    eng.sweep(1)
    J1, pi1 = eng.get_J(), eng.get_pi()
    Jr, pr = c_oracle.sweep_fused(P, J0)                     # EVERY node: the oracle needs a second or two for 2e8 evals
    assert np.array_equal(J1, Jr) and np.array_equal(pi1, pr), (int((J1 != Jr).sum()), int((pi1 != pr).sum()))
    # determinism
    eng.set_J(J0)
    eng.sweep(1)
    assert np.array_equal(eng.get_J(), J1) and np.array_equal(eng.get_pi(), pi1)
    # monotone: J0 <= J0' => T(J0) <= T(J0')
    eng.set_J(J0 + rng.uniform(0, 5, P.N))
    eng.sweep(1)
    assert (eng.get_J() >= J1 - 1e-9).all()
    # shift: T(J0 + c) = T(J0) + c where the minimiser is not the INF branch
    eng.set_J(J0 + 10.0)
    eng.sweep(1)
    Js = eng.get_J()
    finite = (J1 < 299.0) & (Js < 299.0)
    assert finite.mean() > 0.5 and np.abs(Js[finite] - J1[finite] - 10.0).max() < 1e-9
    eng.close()


@pytest.mark.parametrize("name,case", [
    ("cartpole_71", dict(system="CartPole", x_grid_dim=[71, 71, 71, 71], u_grid_dim=[51], xbar=[0.0, float(np.pi), 0.0, 0.0], INF=1000.0)),
    ("twolink_51", dict(system="TwoLinkManipulator", x_grid_dim=[51, 51, 51, 51], u_grid_dim=[21, 21], INF=1000.0)),
    ("dpend_51", dict(CASES["dpend_example"], x_grid_dim=[51, 51, 51, 51], u_grid_dim=[31, 31])),
])
def test_large_4d_grids_sampled_against_oracle(name, case):
    """4-D grids of the BASELINE systems at sizes where G = 1 (one thread scans all actions of its node) and the
    grid spans many blocks per (i0,i1) plane: random node ranges of one backup of a rough J against the C oracle."""
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    eng = Engine(P)
    assert eng.lanes_per_node == 1
    rng = np.random.default_rng(2)
    J0 = rng.uniform(0, 300, P.N)
    eng.set_J(J0)
    st = eng.sweep(1)
    J1, pi1 = eng.get_J(), eng.get_pi()
    plane = P.N // P.dims[0]
    starts = list(rng.integers(0, P.N - 256, 20)) + [0, P.N - 256, plane * (P.dims[0] // 2) - 128]
    for lo in starts:
        Jr, pr = c_oracle.sweep_fused(P, J0, int(lo), int(lo) + 256)
        assert np.array_equal(J1[lo:lo + 256], Jr) and np.array_equal(pi1[lo:lo + 256], pr), (name, int(lo))
    d = J1 - J0
    assert st[0, 0] == J1.max() and st[0, 1] == d.max() and st[0, 2] == d.min()
    eng.close()


def test_edge_cases():
    # smallest legal grid (2 levels per axis), a single action, and a grid where every transition leaves the box
    tiny = dict(system="SinglePendulum", x_grid_dim=[2, 2], u_grid_dim=[1], u_lb=[0.0], u_ub=[0.0], INF=9.0)
    _, grid, cf = build_case(tiny)
    P = problem.extract(grid, cf, 1.0)
    eng = Engine(P)
    eng.eval_terminal_cost()
    eng.sweep(3)
    Jr, pr, _ = c_oracle.run(P, 3)
    assert np.array_equal(eng.get_J(), Jr) and np.array_equal(eng.get_pi(), pr) and (pr == 0).all()
    eng.close()
    allout = dict(system="DoublePendulum", x_grid_dim=[3, 3, 3, 3], u_grid_dim=[2, 2], dt=50.0, INF=123.0,
                  x_lb=[-1, -1, 1, 1], x_ub=[1, 1, 2, 2])
    _, grid, cf = build_case(allout)
    P = problem.extract(grid, cf, 1.0)
    eng = Engine(P)
    eng.eval_terminal_cost()
    st = eng.sweep(1)
    assert (eng.get_J() == 123.0).all() and (eng.get_pi() == 0).all() and st[0, 0] == 123.0
    eng.close()


def test_terminal_cost_kernel_and_policy_tools():
    case = dict(CASES["dpend_example"], S=[2.0, 0.3, 0.0, 1.5])
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    eng = Engine(P)
    eng.eval_terminal_cost()
    ogrid, ocost = oracle_objects(case)
    assert np.array_equal(eng.get_J(), ogrid.terminal(ocost))
    assert np.array_equal(eng.get_J(), c_oracle.terminal(P))
    eng.sweep(4)
    pi = eng.get_pi()
    for k in range(2):
        assert np.array_equal(eng.get_input_from_policy(k), grid.get_input_from_policy(pi, k))
    J = eng.get_J()
    eng.clean_infeasible_set(1.0, 7)
    bad = J > cf.INF - 1.0
    J2, pi2 = eng.get_J(), eng.get_pi()
    assert bad.any() and (J2[bad] == cf.INF).all() and (pi2[bad] == 7).all()
    assert np.array_equal(J2[~bad], J[~bad]) and np.array_equal(pi2[~bad], pi[~bad])
    eng.close()


def test_abi_error_behaviour():
    _, grid, cf = build_case(CASES["pend_51x51x11"])
    eng = Engine(problem.extract(grid, cf, 1.0))
    with pytest.raises(RuntimeError, match="no cost-to-go"):
        eng.sweep(1)
    with pytest.raises(ValueError, match="Grid size does not match"):
        eng.set_J(np.zeros(5))
    with pytest.raises(RuntimeError):
        eng.set_lut(np.zeros(2601 * 11 * 2), np.zeros(2601 * 11))  # not a LUT handle
    eng.eval_terminal_cost()
    with pytest.raises(ValueError):
        eng.get_input_from_policy(3)
    eng.close()
    P = problem.extract(grid, cf, 1.0)
    P.c.dims[1] = 1
    with pytest.raises(ValueError):
        Engine(P)


def test_exact_div_equals_ieee_division():
    """The 3-instruction corrected quotient used for the normalised distance == IEEE division."""
    import ctypes as C
    lib = C.CDLL(_lib.LIB_PATH)
    rng = np.random.default_rng(5)
    n = 1 << 20
    den = np.concatenate([np.resize(np.diff(np.linspace(-2 * np.pi, 2 * np.pi, 1001)), n // 2),
                          rng.uniform(1e-3, 10.0, n - n // 2)])
    a = den * rng.uniform(0, 1, n)
    a[:1000] = den[:1000]
    a[1000:2000] = 0.0
    qf, qi = np.empty(n), np.empty(n)
    rc = lib.pdp_test_exact_div(C.c_void_p(a.ctypes.data), C.c_void_p(den.ctypes.data), C.c_void_p(qf.ctypes.data),
                                C.c_void_p(qi.ctypes.data), C.c_int64(n))
    assert rc == 0
    assert np.array_equal(qi, a / den)
    assert np.array_equal(qf, qi)


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("lanes", [1, 4, 16])
@pytest.mark.parametrize("name", ["pend_51x51x11", "pend_101x101x21", "pend_time_41x61x7", "dpend_example", "twolink_soft", "cartpole_swingup"])
def test_every_lane_split_matches_goldens(name, lanes, generic, monkeypatch):
    """G lanes per node (1: a thread scans all actions; 4 / 16: shuffle argmin with np.argmin's first-index rule),
    and for the pendulum both action loops (the one that exploits an ascending input grid and the
    order-agnostic one): every instantiation must reproduce the reference fixtures."""
    if generic and not name.startswith("pend"):
        pytest.skip("only the pendulum kernel has two action loops")
    monkeypatch.setenv("PYRODP_LANES", str(lanes))
    monkeypatch.setenv("PYRODP_GENERIC", str(generic))
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    eng = Engine(problem.extract(grid, cf, case.get("alpha", 1.0)))
    assert eng.lanes_per_node == lanes
    eng.set_J(gold["J0"])
    k = case["snapshots"][1]
    eng.sweep(k)
    assert np.array_equal(eng.get_J(), gold[f"J_{k}"]) and np.array_equal(eng.get_pi(), gold[f"pi_{k}"])
    eng.close()


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("chunks", [1, 3, 8, 64])
@pytest.mark.parametrize("name", ["pend_301", "cartpole_25", "dpend_21"])
def test_host_array_sweep_is_the_same_backup(name, chunks, pinned, monkeypatch):
    """pdp_sweep_host (host arrays in/out, chunk-pipelined copies) == pdp_set_J + pdp_sweep + getters, for any
    chunking (a chunk's backups read its halo planes, which must have been uploaded before it runs); with
    pinned buffers the pipeline is captured into a CUDA graph and replayed, with pageable ones enqueued directly."""
    monkeypatch.setenv("PYRODP_HOST_CHUNKS", str(chunks))
    if pinned:
        import torch

        def host(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t.numpy()
    else:
        def host(a):
            return a
    case = MID[name]
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    eng = Engine(P)
    J0 = np.random.default_rng(3).uniform(0, 300, P.N)
    eng.set_J(J0)
    st_ref = eng.sweep(1)
    J_ref, pi_ref = eng.get_J(), eng.get_pi()
    eng.set_J(np.zeros(P.N))              # make sure the result really comes from the uploaded array
    Jin, Jout, piout = host(J0.copy()), host(np.empty(P.N)), host(np.empty(P.N, dtype=np.int64))
    for rep in range(3 if pinned else 1):  # replays of the captured graph (both J parities) give the same answer
        Jout[:] = -1.0
        J, pi, st = eng.sweep_host(Jin, Jout, piout)
        assert np.array_equal(J, J_ref) and np.array_equal(pi, pi_ref) and np.array_equal(st, st_ref[0])
    assert np.array_equal(eng.get_J(), J_ref) and np.array_equal(eng.get_J_next(), J0) and np.array_equal(eng.get_pi(), pi_ref)
    J2, pi2, _ = eng.sweep_host(host(J.copy()))   # chained: second backup equals sweeping on
    eng.set_J(J0)
    eng.sweep(2)
    assert np.array_equal(J2, eng.get_J()) and np.array_equal(pi2, eng.get_pi())
    with pytest.raises(ValueError):
        eng.sweep_host(J0[:-1])
    eng.close()
    # slab handles (what each rank of a multi-GPU run holds): full J_next in, the slab's J / pi out, no exchange needed
    n0 = P.dims[0]
    plane = P.N // n0
    for b, e in ((0, n0 // 3), (n0 // 3, 2 * n0 // 3 + 1), (2 * n0 // 3 + 1, n0)):
        slab = Engine(problem.extract(grid, cf, case.get("alpha", 1.0), slab=(b, e)))
        Js, pis, _ = slab.sweep_host(host(J0.copy()), host(np.empty((e - b) * plane)), host(np.empty((e - b) * plane, dtype=np.int64)))
        assert np.array_equal(Js, J_ref[b * plane:e * plane]) and np.array_equal(pis, pi_ref[b * plane:e * plane]), (b, e)
        slab.close()


@pytest.mark.parametrize("name", list(POLICY_CASES))
def test_policy_evaluation_matches_reference_goldens(name):
    """PolicyEvaluatorWithLookUpTable (dynamicprogramming.py:677-752) through the public mirror class: tables built on
    the host as the reference builds them, sweeps in LUT mode with one column per node; bit-exact J."""
    case, gold = POLICY_CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    pe = dynamicprogramming.PolicyEvaluatorWithLookUpTable(LinearFeedback(**case["ctl"]), grid, cf)
    pe.alpha, pe.verbose = case.get("alpha", 1.0), False
    assert np.array_equal(pe.J, gold["J0"])
    k = 0
    for target in case["snapshots"]:
        pe.compute_steps(target - k)
        k = target
        assert np.array_equal(pe.J, gold[f"J_{k}"]) and (pe.pi == 0).all()
    assert np.array_equal(pe.x_next_table, gold["x_next_table"]) and np.array_equal(pe.G, gold["G"])
    # the per-node base class (dynamicprogramming.py:636-672): exactly INF wherever the input is disallowed
    kb = case["snapshots"][1]
    pb = dynamicprogramming.PolicyEvaluator(LinearFeedback(**case["ctl"]), grid, cf)
    pb.alpha, pb.verbose = case.get("alpha", 1.0), False
    pb.compute_steps(kb)
    assert np.array_equal(pb.J, gold[f"Jbase_{kb}"])


# ---- the 4-D range kernel (sweep_mech2.cuh): every variant against the C oracle ----------------------------------
@pytest.mark.parametrize("mode", ["generic", "range"])
@pytest.mark.parametrize("name,case", [
    ("cartpole_41", dict(system="CartPole", x_grid_dim=[31, 33, 41, 43], u_grid_dim=[51], xbar=[0.0, float(np.pi), 0.0, 0.0], INF=1000.0)),
    ("twolink_33", dict(system="TwoLinkManipulator", x_grid_dim=[31, 29, 33, 37], u_grid_dim=[21, 21], INF=1000.0)),
    ("twolink_soft_33", dict(CASES["twolink_soft"], x_grid_dim=[31, 29, 33, 37], u_grid_dim=[9, 7])),
    ("dpend_ex_35", dict(CASES["dpend_example"], x_grid_dim=[33, 31, 35, 37], u_grid_dim=[31, 31])),
    ("dpend_inf_cost", dict(CASES["dpend_example"], x_grid_dim=[21, 23, 25, 27], u_grid_dim=[5, 7], INF=float("inf"))),
])
def test_range_kernel_variants_equal_c_oracle(monkeypatch, name, case, mode):
    """PYRODP_MECH2 pins the 4-D kernel variant: the order-agnostic kernel or the range-skipping kernel.  One lane per node
    (PYRODP_LANES=1), rough J, whole grid compared bit for bit."""
    monkeypatch.setenv("PYRODP_MECH2", mode)
    monkeypatch.setenv("PYRODP_LANES", "1")
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    eng = Engine(P)
    want = {"generic": "sweep_mech2_kernel<", "range": "sweep_mech2_range_kernel<"}[mode]
    assert want in eng.kernel_info and eng.lanes_per_node == 1, eng.kernel_info
    J0 = np.random.default_rng(4).uniform(0, 300, P.N)
    eng.set_J(J0)
    st = eng.sweep(1)
    J1, pi1 = eng.get_J(), eng.get_pi()
    Jr, pr = c_oracle.sweep_fused(P, J0)
    assert np.array_equal(J1, Jr) and np.array_equal(pi1, pr), (name, mode, int((pi1 != pr).sum()))
    d = J1 - J0
    if np.isfinite(J1).all():
        assert st[0, 0] == J1.max() and st[0, 1] == d.max() and st[0, 2] == d.min()
    eng.close()


def _sampled_ranges(P, rng, count, width):
    """Node ranges for sampled checks: random ones plus the first / last nodes, the middle of the grid, and the seams of an
    8-way slab decomposition of axis 0 (the ranges straddle the plane boundary)."""
    plane = P.N // P.dims[0]
    starts = [int(s) for s in rng.integers(0, P.N - width, count)] + [0, P.N - width, plane * (P.dims[0] // 2) - width // 2]
    starts += [plane * (r * P.dims[0] // 8) - width // 2 for r in range(1, 8)]
    return starts


@pytest.mark.parametrize("wl", ["cfg3", "cfg4", "cfg5"])
def test_full_size_4d_configs_sampled_against_oracle(wl):
    """BASELINE configs 3, 4, 5 AT FULL SIZE (TwoLinkManipulator 101^4 x 21^2, CartPole 151^4 x 51, DoublePendulum
    201^4 x 31^2 with the example's bounds): J after two sweeps from h(x) is the input, one more sweep on the GPU, and
    >= 24 + 10 node ranges of 256 nodes (incl. first/last planes and the seams of an 8-rank slab layout) must equal the
    C oracle's backup of the same J_next bit for bit (J and pi)."""
    from bench import WORKLOADS
    case = WORKLOADS[wl]
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    eng = Engine(P)
    assert "range_kernel" in eng.kernel_info, eng.kernel_info
    eng.eval_terminal_cost()
    eng.sweep(2)
    J_next = eng.get_J()              # full host copy: the oracle reads the corners it needs from it
    st = eng.sweep(1)
    rng = np.random.default_rng(7)
    bad = 0
    for lo in _sampled_ranges(P, rng, 24, 256):
        Jr, pr = c_oracle.sweep_fused(P, J_next, lo, lo + 256)
        Jg, pg = eng.get_range("J", lo, 256), eng.get_range("pi", lo, 256)
        assert np.array_equal(eng.get_range("J_next", lo, 256), J_next[lo:lo + 256])
        if not (np.array_equal(Jg, Jr) and np.array_equal(pg, pr)):
            bad += 1
            print(wl, "range", lo, "J mismatches", int((Jg != Jr).sum()), "pi mismatches", int((pg != pr).sum()))
    assert bad == 0
    assert np.isfinite(st).all() and st[0, 0] >= 0.0
    eng.close()


# ---- the unmodified reference on the GPU box (baseline/_ref travels with the snapshot) ---------------------------
from oracle import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference not installed (baseline/_ref)")


@needs_ref
@pytest.mark.parametrize("name", ["pend_51x51x11", "pend_time_41x61x7", "dpend_example", "cartpole_swingup"])
def test_integration_stub_on_the_real_pyro_classes(name):
    """INTEGRATION.md's binding (pyro_b200/pyro_binding.py): a subclass of the REAL pyro DynamicProgramming over the REAL
    GridDynamicSystem / cost function objects, one pdp_sweep_host call per sweep, against the reference's own goldens."""
    from pyro_b200.pyro_binding import bind
    ns = ref_loader.load()
    case, gold = CASES[name], load_golden(name)
    B200 = bind(ns.dynamicprogramming.DynamicProgramming)
    with ref_loader.quiet():
        # the real classes, no look-up tables (lookup=False): the binding never needs them
        _, rgrid, rcf, _ = ref_loader.build_reference(ns, dict(case, x_grid_dim=[3] * len(case["x_grid_dim"])), lut=False)
        sys_ = rgrid.sys
        rgrid = ns.discretizer.GridDynamicSystem(sys_, case["x_grid_dim"], case["u_grid_dim"], case.get("dt", 0.05), False)
        dp = B200(rgrid, rcf)
        dp.alpha = case.get("alpha", 1.0)
        assert isinstance(dp, ns.dynamicprogramming.DynamicProgramming) and np.array_equal(dp.J, gold["J0"])
        k = 0
        for target in case["snapshots"][:3]:
            dp.compute_steps(target - k)          # pyro's own driver loop and finalize_backward_step
            k = target
            assert np.array_equal(dp.J, gold[f"J_{k}"]) and np.array_equal(dp.pi, gold[f"pi_{k}"]), (name, k)
        assert dp.k == k and len(dp.J_list) == k + 1      # the reference's history bookkeeping ran
    dp.close()


@needs_ref
@pytest.mark.parametrize("which", ["InvertedPendulum", "Acrobot"])
def test_subclasses_that_override_the_dynamics_run_in_lut_mode_and_match_the_reference(which):
    """pendulum.InvertedPendulum flips the sign of g, pendulum.Acrobot replaces B.  Both must run on
    tables built by their OWN methods (LUT mode), and equal the reference's DynamicProgrammingWithLookUpTable."""
    ns = ref_loader.load()
    with ref_loader.quiet():
        sys_ = getattr(ns.pendulum, which)()
        dims, udims = ([21, 21], [5]) if which == "InvertedPendulum" else ([5, 5, 5, 5], [3])
        rgrid = ns.discretizer.GridDynamicSystem(sys_, dims, udims, 0.05)
        rcf = ns.costfunction.QuadraticCostFunction.from_sys(sys_)
        rcf.INF = 300
        rdp = ns.dynamicprogramming.DynamicProgrammingWithLookUpTable(rgrid, rcf)
        rdp.compute_steps(6)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(rgrid, rcf)
    dp.verbose = False
    assert dp._engine.problem.system_id == _lib.PDP_SYS_LUT
    dp.compute_steps(6)
    assert np.array_equal(dp.J, rdp.J) and np.array_equal(dp.pi, rdp.pi), which


# ---- single-process multi-part engine (pdp_multi_*): slab + halo logic on ONE GPU, every GPU when there are several ----
def _devices(n_parts):
    import torch
    return [i % torch.cuda.device_count() for i in range(n_parts)]


@pytest.mark.parametrize("n_parts", [2, 3, 5])
@pytest.mark.parametrize("name", ["pend_101x101x21", "pend_time_41x61x7", "dpend_example", "twolink_soft", "cartpole_swingup"])
def test_multi_engine_equals_reference_goldens(name, n_parts):
    """n slab handles driven by one host thread (several per GPU on a 1-GPU box): boundary planes first, peer copies of
    the halo planes under the interior, asymmetric halos (pend_time: 4+5 rows).  J / pi / statistics bit for bit."""
    from pyro_b200.engine import MultiEngine
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    try:
        eng = MultiEngine(P, devices=_devices(n_parts))
    except ValueError as exc:           # the halo does not fit that many slabs of this small grid
        assert "wider than a slab" in str(exc) or "more parts" in str(exc)
        pytest.skip(str(exc))
    eng.eval_terminal_cost()
    assert np.array_equal(eng.get_J(), gold["J0"])
    k = 0
    for target in case["snapshots"][:3]:
        stats = eng.sweep(target - k)
        k = target
        J, pi, Jn = eng.get_J(), eng.get_pi(), eng.get_J_next()
        assert np.array_equal(J, gold[f"J_{k}"]) and np.array_equal(pi, gold[f"pi_{k}"]), (name, n_parts, k)
        d = J - Jn
        assert stats[-1, 0] == J.max() and stats[-1, 1] == d.max() and stats[-1, 2] == d.min()
    assert eng.launch_count >= k * n_parts
    eng.close()


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_multi_engine_random_J_nowait_and_cleaning(monkeypatch, overlap):
    """Mid-size 4-D grid, rough J, the non-blocking enqueue / collect form, clean_infeasible_set with halo refresh and
    get_input_from_policy — against one whole-grid handle."""
    from pyro_b200.engine import MultiEngine
    monkeypatch.setenv("PYRODP_MULTI_OVERLAP", overlap)
    case = dict(system="CartPole", x_grid_dim=[33, 21, 19, 23], u_grid_dim=[9], xbar=[0.0, float(np.pi), 0.0, 0.0], INF=1000.0)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    J0 = np.random.default_rng(7).uniform(0, 300, P.N)
    one = Engine(P)
    one.set_J(J0)
    s_one = one.sweep(3)
    eng = MultiEngine(P, devices=_devices(4))
    eng.set_J(J0)
    eng.sweep_nowait(); eng.sweep_nowait(); eng.sweep_nowait()
    s_multi = eng.collect_stats()
    assert np.array_equal(s_multi, s_one)
    assert np.array_equal(eng.get_J(), one.get_J()) and np.array_equal(eng.get_pi(), one.get_pi())
    assert np.array_equal(eng.get_input_from_policy(0), one.get_input_from_policy(0))
    one.clean_infeasible_set(1.0, 3); eng.clean_infeasible_set(1.0, 3)
    one.sweep(2); eng.sweep(2)
    assert np.array_equal(eng.get_J(), one.get_J()) and np.array_equal(eng.get_pi(), one.get_pi())
    one.close(); eng.close()


def test_planner_uses_every_gpu_without_a_launcher(monkeypatch):
    """DynamicProgrammingWithLookUpTable(grid, cf).compute_steps(k) in a plain script drives all visible GPUs
    (here forced on through PYRODP_MULTI so that a small grid and a 1-GPU box exercise the path too)."""
    monkeypatch.setenv("PYRODP_MULTI", "3")
    case, gold = CASES["dpend_example"], load_golden("dpend_example")
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
    dp.verbose = False
    assert type(dp._engine).__name__ == "MultiEngine" and dp._engine.n_parts == 3
    dp.compute_steps(2)
    assert np.array_equal(dp.J, gold["J_2"]) and np.array_equal(dp.pi, gold["pi_2"])
    dp.alpha = 0.5                       # a parameter change rebuilds the device state and carries J over
    dp.compute_steps(1)
    J3, _ = c_oracle.sweep_fused(problem.extract(grid, cf, 0.5), gold["J_2"])
    assert np.array_equal(dp.J, J3)
    ctl = dp.get_lookup_table_controller()
    assert ctl.c(np.zeros(4), ctl.rbar).shape == (2,)


@pytest.mark.parametrize("name", ["pend_101x101x21", "pend_time_41x61x7", "dpend_example", "twolink_soft", "cartpole_swingup",
                                  "pend_reach_41x41x3", "cartpole_domaincheck"])
def test_device_built_lookup_tables_equal_the_reference_tables(name):
    """The step before the sweep: x_next_table, x_next_isok and G built by pdp_build_tables against the
    samples of the reference's own tables stored in the fixtures (discretizer.py:342-376, dynamicprogramming.py:517-553),
    and the mirror grid's lookup=True path, which uses the device builder on a GPU box."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    eng = Engine(problem.extract(grid, cf, case.get("alpha", 1.0)))
    x_next, x_ok, G = eng.build_tables()
    stride = int(gold["table_stride"])
    assert np.array_equal(x_next[::stride], gold["x_next_sample"])
    assert np.array_equal(x_ok[::stride], gold["x_next_isok_sample"])
    assert np.array_equal(G[::stride], gold["G_sample"])
    lo, cnt = grid.nodes_n // 3, 77                                  # a node range
    xr, okr, Gr = eng.build_tables(lo, cnt)
    assert np.array_equal(xr, x_next[lo:lo + cnt]) and np.array_equal(okr, x_ok[lo:lo + cnt]) and np.array_equal(Gr, G[lo:lo + cnt])
    eng.close()
    _, lgrid, _ = build_case(case, lookup=True)                      # GridDynamicSystem(..., lookup=True) on the mirror
    assert np.array_equal(lgrid.x_next_table, x_next) and np.array_equal(lgrid.x_next_isok, x_ok)
    assert gold["action_isok_sample"].all() and lgrid.action_isok.all()


@pytest.mark.parametrize("tile_rows", ["1", "2", "4", "8", "16"])
@pytest.mark.parametrize("name,case", [
    ("cartpole_ragged", dict(system="CartPole", x_grid_dim=[5, 7, 37, 45], u_grid_dim=[21], xbar=[0.0, float(np.pi), 0.0, 0.0], INF=1000.0)),
    ("dpend_ex_ragged", dict(CASES["dpend_example"], x_grid_dim=[5, 6, 35, 41], u_grid_dim=[9, 11])),
])
def test_range_kernel_block_tiles_equal_c_oracle(monkeypatch, name, case, tile_rows):
    """PYRODP_TILE_ROWS: a block of the range kernel is 128 consecutive nodes (1) or a tile of 2 ... 16 adjacent i2 rows; plane
    sizes that are no multiple of any tile (ragged last tiles in both directions)."""
    monkeypatch.setenv("PYRODP_TILE_ROWS", tile_rows)
    monkeypatch.setenv("PYRODP_LANES", "1")
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    eng = Engine(P)
    assert "range_kernel" in eng.kernel_info
    J0 = np.random.default_rng(12).uniform(0, 300, P.N)
    eng.set_J(J0)
    st = eng.sweep(1)
    Jr, pr = c_oracle.sweep_fused(P, J0)
    assert np.array_equal(eng.get_J(), Jr) and np.array_equal(eng.get_pi(), pr), (name, tile_rows)
    assert st[0, 0] == Jr.max()
    eng.close()
