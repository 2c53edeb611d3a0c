"""Tier-0 oracle: load the UNMODIFIED reference (SherbyRobotics/pyro) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyro_b200/`` may import this module; only
``tests/``, ``oracle/gen_golden.py`` (the golden generator) and documentation scripts use it.

The reference imports matplotlib at module top (pyro/control/controller.py:9,
pyro/planning/discretizer.py:9-12, pyro/planning/dynamicprogramming.py:10-11,
pyro/analysis/graphical.py:10-12) and matplotlib is not installed in this image, so
in-memory ``MagicMock`` modules are injected for it before the import.
Plot methods become no-ops; never pass ``animate_*=True``.

The reference is looked up in ``$PYRO_REF``, then ``/root/reference`` (this container only), then
``baseline/_ref`` — the unmodified package installed by
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``
(git-ignored, but it travels to the GPU box).  Callers must check :func:`available` and skip.
"""
import contextlib
import io
import os
import sys
from unittest import mock

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("PYRO_REF", ""), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")]

_MPL = ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
        "matplotlib.cm", "matplotlib.ticker", "matplotlib.patches", "matplotlib.backends",
        "matplotlib.backends.backend_agg", "matplotlib.figure", "matplotlib.lines",
        "matplotlib.widgets", "mpl_toolkits", "mpl_toolkits.mplot3d",
        "mpl_toolkits.mplot3d.axes3d", "mpl_toolkits.mplot3d.art3d"]


def ref_root():
    for c in _CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "pyro", "planning")):
            return c
    return None


def available():
    return ref_root() is not None


def load():
    """Return a namespace with the reference modules the hot path uses."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference pyro not found (set $PYRO_REF); tier-0 oracle unavailable")
    try:
        import matplotlib  # noqa: F401
    except Exception:
        for name in _MPL:
            sys.modules.setdefault(name, mock.MagicMock(name=name))
    if root not in sys.path:
        sys.path.insert(0, root)
    with contextlib.redirect_stdout(io.StringIO()):
        from pyro.analysis import costfunction
        from pyro.dynamic import cartpole, manipulator, pendulum
        from pyro.planning import discretizer, dynamicprogramming

    class NS:
        pass

    ns = NS()
    ns.costfunction, ns.cartpole, ns.manipulator, ns.pendulum = costfunction, cartpole, manipulator, pendulum
    ns.discretizer, ns.dynamicprogramming = discretizer, dynamicprogramming
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints a line per sweep / table build; silence it in tests."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def build_reference(ns, case, lut=True):
    """Instantiate a parity / benchmark case (plain data, tests/cases.py) on the REAL reference classes:
    (sys, grid_sys, cost_function, dp).  ``lut=False`` builds the grid without look-up tables and returns the base
    class DynamicProgramming (per-pair sys.f calls, dynamicprogramming.py:195-236)."""
    import numpy as np
    cls = {"SinglePendulum": ns.pendulum.SinglePendulum, "DoublePendulum": ns.pendulum.DoublePendulum,
           "TwoLinkManipulator": ns.manipulator.TwoLinkManipulator, "CartPole": ns.cartpole.CartPole}[case["system"]]
    sys_ = cls()
    for key in ("x_lb", "x_ub", "u_lb", "u_ub"):
        if key in case:
            getattr(sys_, key)[:] = case[key]
    for key, val in case.get("sys_params", {}).items():
        setattr(sys_, key, val)
    grid = ns.discretizer.GridDynamicSystem(sys_, case["x_grid_dim"], case["u_grid_dim"], case.get("dt", 0.05), lut)
    from tests.cases import make_cost
    cf = make_cost(ns.costfunction, sys_, case)
    klass = ns.dynamicprogramming.DynamicProgrammingWithLookUpTable if lut else ns.dynamicprogramming.DynamicProgramming
    dp = klass(grid, cf)
    dp.alpha = case.get("alpha", 1.0)
    return sys_, grid, cf, dp
