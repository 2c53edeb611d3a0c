"""pyro_b200 — B200-native value-iteration engine behind pyro's grid-DP API.

Only the hot path of SherbyRobotics/pyro's ``DynamicProgramming`` over ``GridDynamicSystem`` is
here: the Bellman sweep runs as hand-written sm_100a CUDA kernels behind
a C ABI (``include/pyrodp.h``); this package is the thin host-side mirror of the reference's
Python interface for that path.
"""
from . import costfunction, discretizer, dynamicprogramming, systems  # noqa: F401
from .dynamicprogramming import (DynamicProgramming, DynamicProgrammingWithLookUpTable, LookUpTableController,  # noqa: F401
                                 DynamicProgramming2DRectBivariateSpline,
                                 PolicyEvaluator, PolicyEvaluatorWithLookUpTable)
from .discretizer import GridDynamicSystem  # noqa: F401

__version__ = "0.1.0"
