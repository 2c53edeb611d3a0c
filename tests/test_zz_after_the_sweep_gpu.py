"""GPU parity of the closed-loop rollout batches (pdp_rollout, rollout.cuh) through the C ABI and the mirror's public
method, against trajectories of the unmodified reference (`(ctl + sys).compute_trajectory(tf, n, 'euler')`, fixtures
tests/golden/rollout_*.npz written by oracle/gen_golden.py).  Floating point: <= 1e-9 of the trajectory's scale.
(The file sorts after the bit-parity suites on purpose: those are the gate.)"""
import numpy as np
import pytest

from pyro_b200 import dynamicprogramming
from tests.cases import CASES, build_case
from tests.test_kernels_emulated import ROLLOUT_FIXTURES, check_rollout
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ROLLOUT_FIXTURES)
def test_rollout_batches_match_reference_closed_loop_trajectories(name):
    case, gold = CASES[name], load_golden("rollout_" + name)
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
    dp.alpha, dp.verbose = case.get("alpha", 1.0), False
    dp.compute_steps(int(gold["sweeps"]))
    assert np.array_equal(dp.pi, gold["pi"])                       # the policy the reference simulated
    npts, tf = int(gold["npts"]), float(gold["tf"])
    launches = dp._engine.launch_count
    t, x, u = dp.compute_closed_loop_trajectories(gold["x0"], tf, npts)
    assert dp._engine.launch_count == launches + 1                 # one kernel for the whole batch
    assert np.array_equal(t, np.linspace(0, tf, npts))
    check_rollout(x, u, gold)
    # strided output and a large batch of repeated initial states: same trajectories, every thread on its own
    reps = 257
    t7, x7, u7 = dp.compute_closed_loop_trajectories(np.tile(gold["x0"], (reps, 1)), tf, npts, stride=7)
    B = gold["x0"].shape[0]
    assert x7.shape == (reps * B, (npts - 1) // 7 + 1, grid.sys.n) and np.array_equal(t7, t[::7])
    for r in (0, 1, reps - 1):
        assert np.array_equal(x7[r * B:(r + 1) * B], x[:, ::7]) and np.array_equal(u7[r * B:(r + 1) * B], u[:, ::7])


def test_rollout_argument_errors():
    case = CASES["pend_51x51x11"]
    _, grid, cf = build_case(case)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, cf)
    dp.verbose = False
    dp.compute_steps(2)
    with pytest.raises(ValueError):
        dp.compute_closed_loop_trajectories(np.zeros((3, 4)), 1.0, 11)      # wrong state dimension
    with pytest.raises(ValueError):
        dp._engine.rollout(np.zeros(16), np.zeros((1, 2)), 0, 0.1)           # npts < 1
    with pytest.raises(ValueError):
        dp.compute_closed_loop_trajectories(np.zeros((1, 2)), 1.0, 1)       # one point: no step size


# ---- DynamicProgramming2DRectBivariateSpline (pdp_set_interpolant, spline.cuh) ---------------------------------------
from oracle import np_oracle as npo                                       # noqa: E402
from pyro_b200 import problem                                              # noqa: E402
from pyro_b200.engine import Engine                                        # noqa: E402
from tests.test_kernels_emulated import SPLINE_FIXTURES, check_spline_snapshot   # noqa: E402


@pytest.mark.parametrize("name", SPLINE_FIXTURES)
def test_spline_class_matches_reference_spline_class(name):
    """The mirror of DynamicProgramming2DRectBivariateSpline (dynamicprogramming.py:578-614) on the device against J / pi
    snapshots of the unmodified reference class."""
    case, gold = CASES[name], load_golden("spline_" + name)
    _, grid, cf = build_case(case, lookup=True)
    dp = dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid, cf)
    dp.alpha, dp.verbose = case.get("alpha", 1.0), False
    assert np.array_equal(dp.J, gold["J0"])
    k = 0
    for target in gold["snapshots"]:
        dp.compute_steps(int(target) - k)
        k = int(target)
        mism = check_spline_snapshot(dp.J, dp.pi, gold, k)
        print(f"spline {name} k={k}: max |dJ| {np.abs(dp.J - gold[f'J_{k}']).max():.2e}, pi differences on tied nodes {mism}")
    assert "spline" in dp._engine.kernel_info
    assert dp._engine.launch_count == 3 * k                                # two fit kernels + the sweep, per backup


def test_spline_sweep_on_rough_J_equals_scipy_oracle():
    """Mid-size grid, random J_next (the spline overshoots between samples), discount: one backup through the C ABI
    against the oracle's scipy call."""
    case = dict(system="SinglePendulum", x_grid_dim=[151, 203], u_grid_dim=[21], xbar=[-3.14, 0.0], INF=300.0)
    _, grid, cf = build_case(case, lookup=True)
    P = problem.extract(grid, cf, 0.95, force_lut=True)
    x_next, G = dynamicprogramming.build_lookup_tables(grid, cf, exact_inf=False)
    eng = Engine(P)
    eng.set_lut(x_next, G)
    eng.set_interpolant("spline3")
    J0 = np.random.default_rng(8).uniform(0, 300, P.N)
    eng.set_J(J0)
    st = eng.sweep(1)
    J, pi = eng.get_J(), eng.get_pi()
    Jr, pr, gap = npo.spline_sweep(grid.x_level, grid.x_grid_dim, J0, x_next, G, 0.95)
    scale = np.abs(Jr).max()
    assert np.abs(J - Jr).max() <= 1e-9 * scale
    assert not ((pi != pr) & (gap > 1e-8 * scale)).any()
    assert st[-1][0] == J.max()
    eng.set_interpolant("linear")                                          # and back: the RGI sweep of the same handle
    eng.set_J(J0)
    eng.sweep(1)
    Jl, pl = npo.lut_sweep(grid.x_level, grid.x_grid_dim, J0, x_next, G, 0.95)
    assert np.array_equal(eng.get_J(), Jl) and np.array_equal(eng.get_pi(), pl)
    eng.close()


def test_spline_interpolant_argument_errors():
    _, grid, cf = build_case(CASES["pend_51x51x11"])
    eng = Engine(problem.extract(grid, cf, 1.0))                          # fused handle: no tables to interpolate over
    with pytest.raises(NotImplementedError):
        eng.set_interpolant("spline3")
    eng.close()
    _, grid4, cf4 = build_case(CASES["twolink_9"])
    with pytest.raises(NotImplementedError):
        dynamicprogramming.DynamicProgramming2DRectBivariateSpline(grid4, cf4)


# ---- a 3-D example of the reference on the device (table mode, n = 3) ----------------------------------------------------
def test_reference_3d_example_drops_in_with_real_pyro_objects():
    """examples/demos_by_tool/dynamicprogramming/helicopter_tunnel.py (coarse grid) with only the planner class swapped:
    the REAL pyro plant (obstacles in isavalidstate), grid with its look-up tables and cost function from the unmodified
    reference under baseline/_ref -> table-mode handle, sweep_lut_kernel<3> -> J / pi of the reference bit for bit."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("unmodified reference not present (baseline/_ref)")
    from tests.cases import helicopter_tunnel_example
    ns = ref_loader.load()
    from pyro.dynamic import drone
    gold = load_golden("helicopter_tunnel_15x13x11")
    with ref_loader.quiet():
        sys_, grid, qcf = helicopter_tunnel_example(drone, ns.costfunction, ns.discretizer)
    dp = dynamicprogramming.DynamicProgrammingWithLookUpTable(grid, qcf)
    dp.alpha, dp.verbose = float(gold["alpha"]), False
    assert "sweep_lut_kernel" in dp._engine.kernel_info and np.array_equal(dp.J, gold["J0"])
    k = 0
    for target in gold["snapshots"]:
        dp.compute_steps(int(target) - k)
        k = int(target)
        assert np.array_equal(dp.J, gold[f"J_{k}"]) and np.array_equal(dp.pi, gold[f"pi_{k}"]), k
