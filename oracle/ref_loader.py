"""Tier-0 oracle: load the UNMODIFIED reference (SherbyRobotics/pyro) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyro_b200/`` may import this module; only
``tests/``, ``oracle/gen_golden.py`` (the golden generator) and documentation scripts use it.

The reference imports matplotlib at module top (pyro/control/controller.py:9,
pyro/planning/discretizer.py:9-12, pyro/planning/dynamicprogramming.py:10-11,
pyro/analysis/graphical.py:10-12) and matplotlib is not installed in this image, so
in-memory ``MagicMock`` modules are injected for it before the import.
Plot methods become no-ops; never pass ``animate_*=True``.

The reference is looked up in ``$PYRO_REF`` then ``/root/reference``.  It does NOT exist on
the GPU box: callers must check :func:`available` and skip.
"""
import contextlib
import io
import os
import sys
from unittest import mock

_CANDIDATES = [os.environ.get("PYRO_REF", ""), "/root/reference"]

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
