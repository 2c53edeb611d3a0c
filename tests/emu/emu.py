"""ctypes access to tests/emu/libpyrodp_emu.so — the product's fused kernels compiled for and run on the CPU.
TEST INFRASTRUCTURE ONLY (see tests/emu/cuda_emu.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpyrodp_emu.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, f) for f in ("emu_main.cpp", "cuda_emu.h", "build.sh")]
        root = os.path.dirname(os.path.dirname(HERE))
        srcs += [os.path.join(root, "pyro_b200", "csrc", f) for f in ("pyrodp_device.cuh", "sweep_fused.cuh", "table_kernels.cuh", "sweep_mech2.cuh", "mech2_plan.h")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
            subprocess.check_call(["bash", os.path.join(HERE, "build.sh")])
        _lib = C.CDLL(LIB)
        _lib.emu_sweep.restype = C.c_int
        _lib.emu_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.emu_sweep_planes.restype = C.c_int
        _lib.emu_sweep_planes.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4
        _lib.emu_lut_sweep.restype = C.c_int
        _lib.emu_lut_sweep.argtypes = [C.c_void_p] * 7 + [C.c_int]
        _lib.emu_mech2_evals.restype = C.c_longlong
        _lib.emu_mech2_evals.argtypes = [C.c_int]
        _lib.emu_build_tables.restype = C.c_int
        _lib.emu_build_tables.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.emu_spline_sweep.restype = C.c_int
        _lib.emu_spline_sweep.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
        _lib.emu_rollout.restype = C.c_int
        _lib.emu_rollout.argtypes = [C.c_void_p] * 4 + [C.c_longlong, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        _lib.emu_terminal.restype = C.c_int
        _lib.emu_terminal.argtypes = [C.c_void_p] * 3
    return _lib


MECH2_MODES = {None: 0, "generic": 1, "range": 0}


def _mode(force_generic, mech2):
    """Kernel selection code of emu_sweep*: 0 = as the library selects (the 4-D range kernel for one lane per node when the
    action table allows it), 1 = the order-agnostic kernels, 2 = the pendulum kernel's loop nest (PYRODP_PEND_LOOP=2)."""
    return int(force_generic) if force_generic else MECH2_MODES[mech2]


def sweep(problem, J_next, lanes=1, force_generic=False, mech2=None):
    """One backup of the whole grid by the emulated kernel: (J, pi, [j_max, delta_max, delta_min])."""
    J_next = np.ascontiguousarray(J_next, dtype=np.float64)
    J = np.empty(problem.N)
    pi = np.empty(problem.N, dtype=np.int64)
    stats = np.empty(3)
    rc = load().emu_sweep(C.addressof(problem.c), J_next.ctypes.data, J.ctypes.data, pi.ctypes.data, stats.ctypes.data,
                          int(lanes), _mode(force_generic, mech2))
    if rc != 0:
        raise RuntimeError(f"emu_sweep failed ({rc})")
    return J, pi, stats


def lut_sweep(problem, J_next, x_next, G, grid_blocks=24):
    """One LUT-mode backup (generic kernel, or the policy-evaluation kernel when the tables have one column) on a
    persistent grid of `grid_blocks` blocks: (J, pi, stats)."""
    J_next = np.ascontiguousarray(J_next, dtype=np.float64)
    x_next = np.ascontiguousarray(x_next, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64)
    J = np.empty(problem.N)
    pi = np.empty(problem.N, dtype=np.int64)
    stats = np.empty(3)
    rc = load().emu_lut_sweep(C.addressof(problem.c), J_next.ctypes.data, x_next.ctypes.data, G.ctypes.data, J.ctypes.data,
                              pi.ctypes.data, stats.ctypes.data, int(grid_blocks))
    if rc != 0:
        raise RuntimeError(f"emu_lut_sweep failed ({rc})")
    return J, pi, stats


def terminal(problem):
    J = np.empty(problem.N)
    pi = np.full(problem.N, -1, dtype=np.int64)
    rc = load().emu_terminal(C.addressof(problem.c), J.ctypes.data, pi.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"emu_terminal failed ({rc})")
    return J, pi


def sweep_planes(problem, J_next, J, pi, p0, p1, lanes=1, force_generic=False, mech2=None):
    """Backup of axis-0 planes [p0, p1) only, written into the full-size arrays J / pi: what one rank (or one
    boundary / interior launch of it) computes.  Returns the statistics triple of those planes."""
    assert J_next.dtype == np.float64 and J.dtype == np.float64 and pi.dtype == np.int64
    stats = np.empty(3)
    rc = load().emu_sweep_planes(C.addressof(problem.c), J_next.ctypes.data, J.ctypes.data, pi.ctypes.data, stats.ctypes.data,
                                 int(lanes), _mode(force_generic, mech2), int(p0), int(p1))
    if rc != 0:
        raise RuntimeError(f"emu_sweep_planes failed ({rc})")
    return stats


def mech2_evals(reset=True):
    """(node, action) pairs the range kernel actually evaluated since the last reset (test instrumentation)."""
    return int(load().emu_mech2_evals(int(reset)))


def rollout(problem, pi, phys, x0, npts, dt, stride=1):
    """The rollout kernel (pdp_rollout) on the CPU: x (B, n_keep, n), u (B, n_keep, m)."""
    x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
    pi = np.ascontiguousarray(pi, dtype=np.int64)
    phys = np.ascontiguousarray(phys, dtype=np.float64)
    B, keep = x0.shape[0], (npts - 1) // stride + 1
    x = np.empty((keep, problem.n, B))
    u = np.empty((keep, problem.m, B))
    rc = load().emu_rollout(C.addressof(problem.c), pi.ctypes.data, phys.ctypes.data, x0.ctypes.data, B, int(npts), float(dt),
                            int(stride), x.ctypes.data, u.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"emu_rollout failed ({rc})")
    return x.transpose(2, 0, 1), u.transpose(2, 0, 1)


def spline_sweep(problem, J_next, x_next, G, grid_blocks=24):
    """One backup of the bicubic-spline table sweep (pdp_set_interpolant(PDP_INTERP_SPLINE3)): the two fit kernels, then
    sweep_lut_spline_kernel: (J, pi, stats, B-spline coefficients (m0, m1))."""
    J_next = np.ascontiguousarray(J_next, dtype=np.float64)
    x_next = np.ascontiguousarray(x_next, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64)
    J = np.empty(problem.N)
    pi = np.empty(problem.N, dtype=np.int64)
    stats = np.empty(3)
    coef = np.empty(problem.N)
    rc = load().emu_spline_sweep(C.addressof(problem.c), J_next.ctypes.data, x_next.ctypes.data, G.ctypes.data, J.ctypes.data,
                                 pi.ctypes.data, stats.ctypes.data, int(grid_blocks), coef.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"emu_spline_sweep failed ({rc})")
    return J, pi, stats, coef


def build_tables(problem, node_begin=0, count=None):
    """build_tables_kernel (pdp_build_tables) on the CPU: (x_next (K,A,n), x_next_isok (K,A) bool, G (K,A))."""
    count = problem.N - node_begin if count is None else count
    xn = np.empty((count, problem.A, problem.n))
    ok = np.empty((count, problem.A), dtype=np.uint8)
    G = np.empty((count, problem.A))
    rc = load().emu_build_tables(C.addressof(problem.c), int(node_begin), int(count), xn.ctypes.data, ok.ctypes.data, G.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"emu_build_tables failed ({rc})")
    return xn, ok.astype(bool), G
