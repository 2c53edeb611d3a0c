"""Thin object wrapper over the C ABI handle (include/pyrodp.h)."""
import ctypes as C

import numpy as np

from . import _lib
from .problem import Problem


class Engine:
    """One device-resident value-iteration state: J, J_next, pi + the problem tables."""

    def __init__(self, problem: Problem):
        self.lib = _lib.load()
        self.problem = problem
        self.N, self.A, self.n, self.m = problem.N, problem.A, problem.n, problem.m
        h = C.c_void_p()
        _lib.check(self.lib.pdp_create(C.byref(problem.c), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.pdp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(rc, self.h)

    # ---- state ----
    def eval_terminal_cost(self):
        self._ck(self.lib.pdp_eval_terminal_cost(self.h))

    def set_J(self, J):
        J = np.ascontiguousarray(J, dtype=np.float64)
        if J.size != self.N:
            raise ValueError("Grid size does not match data")
        self._ck(self.lib.pdp_set_J(self.h, J.ctypes.data))

    def get_J(self, out=None):
        out = np.empty(self.N, dtype=np.float64) if out is None else out
        self._ck(self.lib.pdp_get_J(self.h, out.ctypes.data))
        return out

    def get_J_next(self, out=None):
        out = np.empty(self.N, dtype=np.float64) if out is None else out
        self._ck(self.lib.pdp_get_J_next(self.h, out.ctypes.data))
        return out

    def get_pi(self, out=None):
        out = np.empty(self.N, dtype=np.int64) if out is None else out
        self._ck(self.lib.pdp_get_pi(self.h, out.ctypes.data))
        return out

    def set_lut(self, x_next, G):
        x_next = np.ascontiguousarray(x_next, dtype=np.float64)
        G = np.ascontiguousarray(G, dtype=np.float64)
        c = self.problem.c
        slab_nodes = (c.slab_end - c.slab_begin) * (self.N // c.dims[0])
        if x_next.size != slab_nodes * self.A * self.n or G.size != slab_nodes * self.A:
            raise ValueError("look-up table size does not match the grid")
        self._ck(self.lib.pdp_set_lut(self.h, x_next.ctypes.data, G.ctypes.data))

    # ---- hot path ----
    def sweep(self, n_sweeps=1):
        """Run n sweeps back to back; returns an (n, 3) array of [j_max, delta_max, delta_min]."""
        stats = np.empty((max(n_sweeps, 0), 3), dtype=np.float64)
        self._ck(self.lib.pdp_sweep(self.h, int(n_sweeps), stats.ctypes.data))
        return stats

    def sweep_async(self):
        self._ck(self.lib.pdp_sweep_async(self.h))

    def commit_sweep(self):
        self._ck(self.lib.pdp_commit_sweep(self.h))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.pdp_set_stream(self.h, C.c_void_p(cuda_stream)))

    def device_buffers(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        self._ck(self.lib.pdp_device_buffers(self.h, *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)  # J_cur, J_new, pi, stats

    # ---- after the sweep ----
    def get_input_from_policy(self, k):
        out = np.empty(self.N, dtype=np.float64)
        self._ck(self.lib.pdp_get_input_from_policy(self.h, int(k), out.ctypes.data))
        return out

    def clean_infeasible_set(self, tol, default_action):
        self._ck(self.lib.pdp_clean_infeasible_set(self.h, float(tol), int(default_action)))

    @property
    def nodes_padded(self):
        return int(self.lib.pdp_nodes_padded(self.h))

    @property
    def launch_count(self):
        return int(self.lib.pdp_launch_count(self.h))

    @property
    def last_sweep_ms(self):
        return float(self.lib.pdp_last_sweep_ms(self.h))
