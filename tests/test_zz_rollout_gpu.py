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
