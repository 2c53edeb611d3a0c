"""ctypes wrapper of oracle/dp_oracle.c (the plain-C restatement).  TEST INFRASTRUCTURE ONLY.

Takes the same ``pdp_problem`` descriptor the CUDA library takes (include/pyrodp.h), so the C
restatement and the kernels are fed byte-identical inputs.
"""
import ctypes as C
import os

import numpy as np

from . import np_oracle


def _lib():
    lib = np_oracle.clib()
    lib.orc_sweep_fused.restype = C.c_int
    lib.orc_sweep_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_sweep_lut.restype = C.c_int
    lib.orc_sweep_lut.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_terminal.restype = C.c_int
    lib.orc_terminal.argtypes = [C.c_void_p, C.c_void_p]
    return lib


def n_threads_default():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def sweep_fused(problem, J_next, lo=0, hi=None, n_threads=None):
    """One backup of nodes lo:hi with on-the-fly dynamics.  problem: pyro_b200.problem.Problem."""
    hi = problem.N if hi is None else hi
    J_next = np.ascontiguousarray(J_next, dtype=np.float64)
    J = np.empty(hi - lo)
    pi = np.empty(hi - lo, dtype=np.int64)
    rc = _lib().orc_sweep_fused(C.addressof(problem.c), J_next.ctypes.data, lo, hi, J.ctypes.data, pi.ctypes.data,
                                n_threads or n_threads_default())
    if rc != 0:
        raise RuntimeError("orc_sweep_fused failed (LUT-mode problem?)")
    return J, pi


def sweep_lut(problem, J_next, x_next, G, lo=0, hi=None, n_threads=None):
    hi = problem.N if hi is None else hi
    J_next = np.ascontiguousarray(J_next, dtype=np.float64)
    x_next = np.ascontiguousarray(x_next, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64)
    J = np.empty(hi - lo)
    pi = np.empty(hi - lo, dtype=np.int64)
    _lib().orc_sweep_lut(C.addressof(problem.c), J_next.ctypes.data, x_next.ctypes.data, G.ctypes.data, lo, hi,
                         J.ctypes.data, pi.ctypes.data, n_threads or n_threads_default())
    return J, pi


def terminal(problem):
    J = np.empty(problem.N)
    _lib().orc_terminal(C.addressof(problem.c), J.ctypes.data)
    return J


def run(problem, n_sweeps, J0=None, n_threads=None):
    J = terminal(problem) if J0 is None else np.array(J0, float)
    pi = np.zeros(problem.N, dtype=np.int64)
    stats = []
    for _ in range(n_sweeps):
        Jn = J
        J, pi = sweep_fused(problem, Jn, n_threads=n_threads)
        d = J - Jn
        stats.append((J.max(), d.max(), d.min()))
    return J, pi, np.array(stats)
