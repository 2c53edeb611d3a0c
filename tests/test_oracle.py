"""The oracle against the reference: golden fixtures (tier-0 outputs) and, when importable, live runs."""
import itertools
import json

import numpy as np
import pytest

from oracle import c_oracle, np_oracle as npo, ref_loader
from pyro_b200 import problem
from tests.cases import CASES, POLICY_CASES, LinearFeedback, build_case, oracle_objects
from tests.conftest import load_golden


@pytest.mark.parametrize("name", list(CASES))
def test_numpy_oracle_matches_reference_goldens(name):
    """Tier-1 NumPy restatement == reference: tables, then J / pi after every recorded sweep count."""
    case, gold = CASES[name], load_golden(name)
    assert json.loads(str(gold["case"])) == json.loads(json.dumps(case)), "golden file is stale; rerun oracle/gen_golden.py"
    grid, cost = oracle_objects(case)
    stride = int(gold["table_stride"])
    # tables (discretizer.py:342-376, dynamicprogramming.py:517-553) on the sampled nodes
    x_next, x_ok, a_ok, G = grid.tables(cost)
    assert np.array_equal(x_next[::stride], gold["x_next_sample"])
    assert np.array_equal(x_ok[::stride], gold["x_next_isok_sample"])
    assert np.array_equal(a_ok[::stride], gold["action_isok_sample"])
    assert np.array_equal(G[::stride], gold["G_sample"])
    # sweeps
    J = grid.terminal(cost)
    assert np.array_equal(J, gold["J0"])
    k = 0
    alpha = case.get("alpha", 1.0)
    for target in case["snapshots"]:
        for _ in range(target - k):
            J, pi = grid.sweep(J, cost, alpha)
        k = target
        assert np.array_equal(J, gold[f"J_{k}"]), f"J differs after {k} sweeps"
        assert np.array_equal(pi, gold[f"pi_{k}"]), f"pi differs after {k} sweeps"


@pytest.mark.parametrize("name", list(CASES))
def test_c_oracle_matches_reference_goldens(name):
    """Plain-C restatement (fed by the product's descriptor) == reference, bit for bit."""
    case, gold = CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, case.get("alpha", 1.0))
    J = c_oracle.terminal(P)
    assert np.array_equal(J, gold["J0"])
    k = 0
    for target in case["snapshots"]:
        for _ in range(target - k):
            J, pi = c_oracle.sweep_fused(P, J)
        k = target
        assert np.array_equal(J, gold[f"J_{k}"])
        assert np.array_equal(pi, gold[f"pi_{k}"])


def test_c_lut_sweep_matches_goldens():
    case, gold = CASES["pend_51x51x11"], load_golden("pend_51x51x11")
    _, grid, cf = build_case(case)
    P = problem.extract(grid, cf, 1.0)
    J, pi = c_oracle.sweep_lut(P, gold["J0"], gold["x_next_sample"], gold["G_sample"])
    assert np.array_equal(J, gold["J_1"]) and np.array_equal(pi, gold["pi_1"])


@pytest.mark.parametrize("n", [2, 3, 4])
def test_rgi_restatement_matches_scipy(n):
    """rgi_linear == scipy RegularGridInterpolator(linear, bounds_error=False, fill_value=0), bitwise,
    on random points, exact level hits, both bounds, and just-outside points."""
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(n)
    dims = [7, 5, 6, 4][:n]
    levels = [np.linspace(-1.0 - d, 2.0 + 0.5 * d, dims[d]) for d in range(n)]
    values = rng.uniform(0, 300, dims)
    pts = [rng.uniform(-1.2, 1.2, (4000, n)) * np.array([lv[-1] - lv[0] for lv in levels]) / 2
           + np.array([(lv[-1] + lv[0]) / 2 for lv in levels])]
    special = [np.concatenate([lv, [np.nextafter(lv[0], -np.inf), np.nextafter(lv[-1], np.inf),
                                    np.nextafter(lv[0], np.inf), np.nextafter(lv[-1], -np.inf)]]) for lv in levels]
    pts.append(np.array(list(itertools.islice(itertools.product(*special), 0, 20000))))
    xi = np.concatenate(pts)
    want = RegularGridInterpolator(tuple(levels), values, "linear", False, 0)(xi)
    got = npo.rgi_linear(levels, values, xi)
    assert np.array_equal(got, want)


def test_scipy_and_own_rgi_give_identical_sweeps():
    grid, cost = oracle_objects(CASES["dpend_example"])
    J0 = np.random.default_rng(0).uniform(0, 300, grid.N)
    a = grid.sweep(J0, cost, 1.0, use_scipy=False)
    b = grid.sweep(J0, cost, 1.0, use_scipy=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_survey_anchor_values():
    """Sanity anchors measured on the reference during the survey."""
    gold = load_golden("pend_51x51x11")
    grid, cost = oracle_objects(CASES["pend_51x51x11"])
    J = grid.terminal(cost)
    for _ in range(5):
        J, _ = grid.sweep(J, cost)
    assert abs(J.max() - 319.38) < 0.01
    x_next, x_ok, a_ok, _ = grid.tables(cost)
    assert a_ok.all()
    assert abs((~x_ok).mean() - 0.0461) < 1e-3
    assert gold["J_100"].shape == (2601,)


@pytest.mark.skipif(not ref_loader.available(), reason="reference not present (GPU box)")
def test_live_reference_base_class_equals_oracle():
    """The literal per-pair formulation (dynamicprogramming.py:195-236) run live == the oracle."""
    ns = ref_loader.load()
    case = dict(system="SinglePendulum", x_grid_dim=[15, 13], u_grid_dim=[5], xbar=[-3.14, 0.0], INF=300.0)
    with ref_loader.quiet():
        s = ns.pendulum.SinglePendulum()
        g = ns.discretizer.GridDynamicSystem(s, case["x_grid_dim"], case["u_grid_dim"], lookup=False)
        cf = ns.costfunction.QuadraticCostFunction.from_sys(s)
        cf.xbar, cf.INF = np.array(case["xbar"]), case["INF"]
        dp = ns.dynamicprogramming.DynamicProgramming(g, cf)
        dp.compute_steps(3)
    grid, cost = oracle_objects(case)
    J, pi, _ = grid.run(cost, 3)
    assert np.array_equal(J, dp.J) and np.array_equal(pi, dp.pi)


@pytest.mark.parametrize("name", list(POLICY_CASES))
def test_policy_evaluation_oracle_and_mirror_tables_match_reference(name):
    """Policy evaluation (dynamicprogramming.py:677-752): the NumPy restatement J = G + alpha*RGI(J_next)(x_next)
    reproduces the reference's snapshots from the reference's tables, and the mirror's compute_lookuptable
    (host loop, same calls) rebuilds those tables bit for bit."""
    case, gold = POLICY_CASES[name], load_golden(name)
    _, grid, cf = build_case(case)
    J = gold["J0"].copy()
    k = 0
    for target in case["snapshots"]:
        for _ in range(target - k):
            J, pi = npo.lut_sweep(grid.x_level, case["x_grid_dim"], J, gold["x_next_table"][:, None, :], gold["G"][:, None],
                                  case.get("alpha", 1.0), use_scipy=False)
            assert (pi == 0).all()
        k = target
        assert np.array_equal(J, gold[f"J_{k}"])
    from pyro_b200 import dynamicprogramming as dpm

    class NoEngine:  # host-side table build only: no device in the CPU suite
        def __init__(self, N):
            self.N, self.problem = N, type("P", (), {"system_id": 0})()
        def set_J(self, J): pass
        def get_J(self): return np.zeros(self.N)
        def close(self): pass
    pe = dpm.PolicyEvaluatorWithLookUpTable(LinearFeedback(**case["ctl"]), grid, cf,
                                            engine_factory=lambda dp, P: NoEngine(grid.nodes_n))
    pe.compute_lookuptable()
    assert np.array_equal(pe.x_next_table, gold["x_next_table"]) and np.array_equal(pe.G, gold["G"])


@pytest.mark.parametrize("name", ["pend_51x51x11", "pend_101x101x21", "cartpole_swingup"])
def test_boundary_audit_near_the_domain_bounds(name):
    """SURVEY.md section 7, hard part 1: pi hinges on the in/out-of-box classification of x_next within an ulp of
    the bounds.  Count the sampled pairs that land exactly on / within 4 ulp of a bound and check that the
    restatement classifies every one of them as the reference did (strict compares: ON the bound is inside)."""
    case, gold = CASES[name], load_golden(name)
    grid, cost = oracle_objects(case)
    stride = int(gold["table_stride"])
    x_next, x_ok, _, _ = grid.tables(cost)
    x_next, x_ok = x_next[::stride], x_ok[::stride]
    lb, ub = np.asarray(grid.spec.x_lb, float), np.asarray(grid.spec.x_ub, float)
    near = np.zeros(x_next.shape[:2], dtype=bool)
    on = np.zeros_like(near)
    for d in range(x_next.shape[2]):
        for b in (lb[d], ub[d]):
            dist = np.abs(x_next[..., d] - b)
            near |= dist <= 4 * np.spacing(abs(b))
            on |= x_next[..., d] == b
    print(f"{name}: {int(on.sum())} pairs exactly on a bound, {int(near.sum())} within 4 ulp, of {near.size}")
    assert np.array_equal(x_ok[near], gold["x_next_isok_sample"][near])
    if name == "pend_51x51x11":
        assert on.sum() > 0                      # e.g. dq = 0 nodes on the q boundary
    inside_otherwise = np.all((x_next >= lb) & (x_next <= ub), axis=-1)
    assert (x_ok[on] == inside_otherwise[on]).all()   # a pair ON one bound is valid unless another axis is outside
