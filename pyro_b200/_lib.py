"""ctypes binding of libpyrodp.so (the C ABI in include/pyrodp.h).

There is deliberately NO fallback: if the shared library is missing, or no CUDA device is
usable, every entry point raises.  The library is built in-tree by ``__graft_entry__.build()``
(nvcc, sm_100a) as ``pyro_b200/libpyrodp.so``.
"""
import ctypes as C
import os

PDP_ABI_VERSION = 1
PDP_MAX_N, PDP_MAX_M = 4, 2
PDP_OK, PDP_EINVAL, PDP_ENOTSUP, PDP_ECUDA, PDP_ESTATE = 0, -1, -2, -3, -4
PDP_SYS_LUT, PDP_SYS_PENDULUM, PDP_SYS_TWOLINK, PDP_SYS_CARTPOLE = 0, 1, 2, 3
PDP_COST_QUADRATIC, PDP_COST_TIME, PDP_COST_REACH = 1, 2, 3
PDP_INTERP_LINEAR, PDP_INTERP_SPLINE3 = 0, 1

_dp = C.POINTER(C.c_double)


class pdp_problem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n", C.c_int32), ("m", C.c_int32),
        ("system_id", C.c_int32), ("cost_id", C.c_int32), ("ontarget_check", C.c_int32),
        ("slab_begin", C.c_int32), ("slab_end", C.c_int32),
        ("alloc_planes", C.c_int32), ("reserved0", C.c_int32),
        ("dims", C.c_int32 * PDP_MAX_N), ("udims", C.c_int32 * PDP_MAX_M),
        ("x_level", _dp * PDP_MAX_N), ("u_level", _dp * PDP_MAX_M),
        ("x_lb", C.c_double * PDP_MAX_N), ("x_ub", C.c_double * PDP_MAX_N),
        ("dt", C.c_double), ("alpha", C.c_double), ("INF", C.c_double), ("EPS", C.c_double),
        ("Q", C.c_double * 16), ("S", C.c_double * 16),
        ("xbar", C.c_double * PDP_MAX_N), ("sys_par", C.c_double * 8),
        ("sys_tab", _dp * 4), ("sys_tab_len", C.c_int64 * 4),
        ("bu", _dp), ("gu", _dp), ("act_ok", C.POINTER(C.c_uint8)),
    ]


class pdp_stats(C.Structure):
    _fields_ = [("j_max", C.c_double), ("delta_max", C.c_double), ("delta_min", C.c_double)]


# PYRODP_LIB: development hook to A/B an alternative build of the SAME library (never a fallback)
LIB_PATH = os.environ.get("PYRODP_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpyrodp.so")

# name -> (restype, argtypes); exactly the symbols include/pyrodp.h declares
SIGNATURES = {
    "pdp_abi_version": (C.c_int, []),
    "pdp_create": (C.c_int, [C.POINTER(pdp_problem), C.POINTER(C.c_void_p)]),
    "pdp_destroy": (C.c_int, [C.c_void_p]),
    "pdp_last_error": (C.c_char_p, [C.c_void_p]),
    "pdp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_eval_terminal_cost": (C.c_int, [C.c_void_p]),
    "pdp_set_J": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_get_J": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_get_J_next": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_get_pi": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_get_range": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]),
    "pdp_kernel_info": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "pdp_sweep": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pdp_sweep_enqueue": (C.c_int, [C.c_void_p]),
    "pdp_sweep_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "pdp_sweep_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdp_sweep_host_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "pdp_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    "pdp_exchange_current": (C.c_int, [C.c_void_p]),
    "pdp_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_peer_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdp_set_lut": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdp_build_tables": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdp_get_input_from_policy": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pdp_set_interpolant": (C.c_int, [C.c_void_p, C.c_int32]),
    "pdp_rollout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "pdp_clean_infeasible_set": (C.c_int, [C.c_void_p, C.c_double, C.c_int64]),
    "pdp_sweep_async": (C.c_int, [C.c_void_p]),
    "pdp_sweep_planes_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "pdp_commit_sweep": (C.c_int, [C.c_void_p]),
    "pdp_read_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_slab_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "pdp_compute_halo": (C.c_int, [C.POINTER(pdp_problem), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pdp_device_buffers": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4),
    "pdp_device_count": (C.c_int, []),
    "pdp_multi_create": (C.c_int, [C.POINTER(pdp_problem), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]),
    "pdp_multi_destroy": (C.c_int, [C.c_void_p]),
    "pdp_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "pdp_multi_parts": (C.c_int32, [C.c_void_p]),
    "pdp_multi_part": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "pdp_multi_part_device": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pdp_multi_launch_count": (C.c_int64, [C.c_void_p]),
    "pdp_multi_eval_terminal_cost": (C.c_int, [C.c_void_p]),
    "pdp_multi_set_J": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pdp_multi_get": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pdp_multi_sweep": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pdp_multi_sweep_enqueue": (C.c_int, [C.c_void_p]),
    "pdp_multi_sweep_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "pdp_multi_get_input_from_policy": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pdp_multi_clean_infeasible_set": (C.c_int, [C.c_void_p, C.c_double, C.c_int64]),
    "pdp_nodes": (C.c_int64, [C.c_void_p]),
    "pdp_nodes_padded": (C.c_int64, [C.c_void_p]),
    "pdp_actions": (C.c_int64, [C.c_void_p]),
    "pdp_launch_count": (C.c_int64, [C.c_void_p]),
    "pdp_last_sweep_ms": (C.c_double, [C.c_void_p]),
}

_lib = None


def load():
    """Load libpyrodp.so once; raise (never fall back) if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "pyro_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.pdp_abi_version() != PDP_ABI_VERSION:
            raise RuntimeError("libpyrodp.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def last_error(handle=None):
    msg = load().pdp_last_error(handle)
    return msg.decode() if msg else ""


def check(rc, handle=None):
    """Map C status codes onto the reference's exception conventions."""
    if rc == PDP_OK:
        return
    msg = last_error(handle) or last_error(None) or f"pdp error {rc}"
    if rc == PDP_EINVAL:
        raise ValueError(msg)
    if rc == PDP_ENOTSUP:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
