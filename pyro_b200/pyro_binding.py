"""The binding a pyro maintainer would add (as ``pyro/planning/dynamicprogramming_b200.py``): a subclass of the REAL
``pyro.planning.dynamicprogramming.DynamicProgramming`` whose sweep runs in ``libpyrodp.so``.

pyro has no FFI; its extension point is overriding the three per-sweep hooks (dynamicprogramming.py:175, :195, :240),
the way the reference ships its own variants (:505, :578, :623).  This module is that override written against the C ABI
of ``include/pyrodp.h`` only — ``pdp_create`` + ONE ``pdp_sweep_host`` call per sweep in the reference's own calling
convention (J_next a host array in, J and pi host arrays out) — so everything else (constructor, ``compute_steps``,
``solve_bellman_equation``, history lists, ``finalize_backward_step``'s print, plots, ``save_latest``) is pyro's
unmodified code.  ``tests/test_parity_gpu.py::test_integration_stub_on_the_real_pyro_classes`` runs it on a B200 against
the reference-generated goldens.  (``pyro_b200.dynamicprogramming.DynamicProgramming`` is the production planner: it keeps
J and pi on the device between sweeps and needs no O(N) host objects.)

    from pyro.planning import dynamicprogramming
    from pyro_b200.pyro_binding import bind
    DynamicProgrammingB200 = bind(dynamicprogramming.DynamicProgramming)
    dp = DynamicProgrammingB200(grid_sys, cost_function)      # grid_sys may be built with lookup=False: no tables needed
    dp.solve_bellman_equation(tol=0.1)
"""
import ctypes as C

import numpy as np

from . import _lib
from .problem import extract


def bind(DynamicProgramming):
    """Return the B200 subclass of the given (real) pyro DynamicProgramming class."""

    class DynamicProgrammingB200(DynamicProgramming):
        """Same constructor, same attributes; the Bellman backup runs on the GPU."""

        _h = None

        def initialize_backward_step(self):
            self.k = self.k + 1
            self.t = self.t - self.grid_sys.dt
            if self._h is None:                     # lazily: users set cf / alpha after __init__
                self._lib = _lib.load()             # raises if libpyrodp.so is absent: no CPU fallback
                self._p = extract(self.grid_sys, self.cf, self.alpha, self.interpol_method)
                if self._p.system_id == _lib.PDP_SYS_LUT:
                    raise NotImplementedError("this minimal binding covers the fused systems; pyro_b200.dynamicprogramming "
                                              "handles arbitrary systems through pdp_set_lut")
                self._h = C.c_void_p()
                _lib.check(self._lib.pdp_create(C.byref(self._p.c), C.byref(self._h)))
            self.J_next = self.J

        def compute_backward_step(self):
            # one call: upload J_next, backup on the device, download J and pi (pipelined over plane chunks)
            J_next = np.ascontiguousarray(self.J_next, dtype=np.float64)
            self.J = np.empty(self.grid_sys.nodes_n)
            self.pi = np.empty(self.grid_sys.nodes_n, dtype=np.int64)
            stats = np.empty(3)
            _lib.check(self._lib.pdp_sweep_host(self._h, J_next.ctypes.data, self.J.ctypes.data, self.pi.ctypes.data,
                                                stats.ctypes.data), self._h)
        # finalize_backward_step is inherited unchanged (prints, history, returns max|dJ|)

        def close(self):
            if self._h is not None:
                self._lib.pdp_destroy(self._h)
                self._h = None

        def __del__(self):
            try:
                self.close()
            except Exception:
                pass

    return DynamicProgrammingB200
