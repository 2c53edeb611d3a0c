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
        lay = (C.c_int32 * 8)()
        self._ck(self.lib.pdp_slab_layout(self.h, lay))
        (self.slab_begin, self.slab_end, self.alloc_begin, self.alloc_end,
         self.halo_lo, self.halo_hi, self.n0, self.lanes_per_node) = (int(v) for v in lay)
        self.plane = self.N // self.n0
        self.slab_nodes = (self.slab_end - self.slab_begin) * self.plane

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

    # getters return this handle's slab (the whole grid on a single GPU)
    def _out(self, out, dtype):
        if out is None:
            return np.empty(self.slab_nodes, dtype=dtype)
        if out.size != self.slab_nodes or out.dtype != dtype or not out.flags.c_contiguous:
            raise ValueError("output buffer does not match the slab size / dtype")
        return out

    def get_J(self, out=None):
        out = self._out(out, np.float64)
        self._ck(self.lib.pdp_get_J(self.h, out.ctypes.data))
        return out

    def get_J_next(self, out=None):
        out = self._out(out, np.float64)
        self._ck(self.lib.pdp_get_J_next(self.h, out.ctypes.data))
        return out

    def get_pi(self, out=None):
        out = self._out(out, np.int64)
        self._ck(self.lib.pdp_get_pi(self.h, out.ctypes.data))
        return out

    def get_range(self, which, node_begin, count, out=None):
        """J ('J'), J_next ('J_next') or pi ('pi') of the global nodes [node_begin, node_begin + count): J / J_next anywhere
        in the planes this handle holds (halo included), pi inside its slab."""
        code = {"J": 0, "J_next": 1, "pi": 2}[which]
        dtype = np.int64 if code == 2 else np.float64
        if out is None:
            out = np.empty(int(count), dtype=dtype)
        elif out.size != count or out.dtype != dtype or not out.flags.c_contiguous:
            raise ValueError("output buffer does not match the range / dtype")
        self._ck(self.lib.pdp_get_range(self.h, code, int(node_begin), int(count), out.ctypes.data))
        return out

    @property
    def kernel_info(self):
        buf = C.create_string_buffer(160)
        self._ck(self.lib.pdp_kernel_info(self.h, buf, 160))
        return buf.value.decode()

    def build_tables(self, node_begin=0, count=None, x_next=True, x_ok=True, G=True):
        """The reference's dense tables of a node range, built on the device (fused systems): (x_next (K,A,n) or None,
        x_next_isok (K,A) bool or None, G (K,A) or None)."""
        count = self.N - node_begin if count is None else count
        xn = np.empty((count, self.A, self.n)) if x_next else None
        ok = np.empty((count, self.A), dtype=np.uint8) if x_ok else None
        Gt = np.empty((count, self.A)) if G else None
        ptr = lambda a: a.ctypes.data if a is not None else None
        self._ck(self.lib.pdp_build_tables(self.h, int(node_begin), int(count), ptr(xn), ptr(ok), ptr(Gt)))
        return xn, (ok.astype(bool) if ok is not None else None), Gt

    def set_lut(self, x_next, G):
        x_next = np.ascontiguousarray(x_next, dtype=np.float64)
        G = np.ascontiguousarray(G, dtype=np.float64)
        if x_next.size != self.slab_nodes * self.A * self.n or G.size != self.slab_nodes * self.A:
            raise ValueError("look-up table size does not match the grid")
        self._ck(self.lib.pdp_set_lut(self.h, x_next.ctypes.data, G.ctypes.data))

    # ---- hot path ----
    def sweep(self, n_sweeps=1):
        """Run n sweeps back to back; returns an (n, 3) array of [j_max, delta_max, delta_min]."""
        stats = np.empty((max(n_sweeps, 0), 3), dtype=np.float64)
        self._ck(self.lib.pdp_sweep(self.h, int(n_sweeps), stats.ctypes.data))
        return stats

    def sweep_host(self, J_next, J_out=None, pi_out=None):
        """One sweep with host arrays on both sides (upload, backup and download pipelined over plane
        chunks): returns (J, pi, [j_max, delta_max, delta_min]).  Pinned buffers overlap the copies."""
        J_next = np.ascontiguousarray(J_next, dtype=np.float64)
        if J_next.size != self.N:
            raise ValueError("Grid size does not match data")
        J_out, pi_out = self._out(J_out, np.float64), self._out(pi_out, np.int64)
        stats = np.empty(3, dtype=np.float64)
        self._ck(self.lib.pdp_sweep_host(self.h, J_next.ctypes.data, J_out.ctypes.data, pi_out.ctypes.data, stats.ctypes.data))
        return J_out, pi_out, stats

    def sweep_host_local(self, J_held, J_out=None, pi_out=None):
        """sweep_host for a rank of a sharded run: J_held holds only the planes [alloc_begin, alloc_end) of J_next."""
        J_held = np.ascontiguousarray(J_held, dtype=np.float64)
        if J_held.size != (self.alloc_end - self.alloc_begin) * self.plane:
            raise ValueError("J_held must hold exactly the planes this handle keeps (slab + halo)")
        J_out, pi_out = self._out(J_out, np.float64), self._out(pi_out, np.int64)
        stats = np.empty(3, dtype=np.float64)
        self._ck(self.lib.pdp_sweep_host_local(self.h, J_held.ctypes.data, J_out.ctypes.data, pi_out.ctypes.data, stats.ctypes.data))
        return J_out, pi_out, stats

    def sweep_nowait(self):
        """Enqueue one sweep (+ exchange when a communicator is attached) without blocking."""
        self._ck(self.lib.pdp_sweep_enqueue(self.h))
        self._enqueued = getattr(self, "_enqueued", 0) + 1

    def collect_stats(self):
        """Wait for the enqueued sweeps; (k, 3) array of [j_max, delta_max, delta_min], reduced over ranks."""
        n = getattr(self, "_enqueued", 0)
        out = np.empty((max(n, 1), 3), dtype=np.float64)
        got = C.c_int32(0)
        self._ck(self.lib.pdp_sweep_collect(self.h, out.ctypes.data, n, C.byref(got)))
        self._enqueued = 0
        return out[:got.value]

    # ---- multi-GPU: native NCCL exchange inside the library ----
    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        self._ck(self.lib.pdp_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, rank, world, unique_id, mode, overlap=True):
        """mode: 'halo' (send/recv with ranks r-1 / r+1) or 'allgather' (whole slabs)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.lib.pdp_comm_init(self.h, int(rank), int(world), buf, {"halo": 1, "allgather": 2}[mode], int(bool(overlap))))

    def peer_export(self):
        buf = C.create_string_buffer(200)
        self._ck(self.lib.pdp_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, lower, upper):
        """lower / upper: peer_export() of ranks r-1 / r+1 (None at the ends)."""
        lo = C.create_string_buffer(bytes(lower), 200) if lower is not None else None
        hi = C.create_string_buffer(bytes(upper), 200) if upper is not None else None
        self._ck(self.lib.pdp_peer_attach(self.h, lo, hi))

    def exchange_current(self):
        self._ck(self.lib.pdp_exchange_current(self.h))

    def sweep_async(self):
        self._ck(self.lib.pdp_sweep_async(self.h))

    def sweep_planes_async(self, plane_begin, plane_end, stat_set=0):
        self._ck(self.lib.pdp_sweep_planes_async(self.h, int(plane_begin), int(plane_end), int(stat_set)))

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
        out = np.empty(self.slab_nodes, dtype=np.float64)
        self._ck(self.lib.pdp_get_input_from_policy(self.h, int(k), out.ctypes.data))
        return out

    def clean_infeasible_set(self, tol, default_action):
        self._ck(self.lib.pdp_clean_infeasible_set(self.h, float(tol), int(default_action)))

    def set_interpolant(self, which):
        """'linear' (RegularGridInterpolator, the default) or 'spline3' (RectBivariateSpline kx = ky = 3: table-mode handles
        of 2-D grids, dynamicprogramming.py:578-614)."""
        code = {"linear": _lib.PDP_INTERP_LINEAR, "spline3": _lib.PDP_INTERP_SPLINE3}[which]
        self._ck(self.lib.pdp_set_interpolant(self.h, code))

    def rollout(self, phys, x0, npts, dt, stride=1, with_inputs=True):
        """B closed-loop Euler trajectories under the current policy (pdp_rollout): x (B, n_keep, n), u (B, n_keep, m)."""
        x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)))
        phys = np.ascontiguousarray(phys, dtype=np.float64)
        n, m = int(self.problem.n), int(self.problem.m)
        if x0.shape[1] != n or phys.size != 16:
            raise ValueError("rollout: x0 must be (B, n) and phys 16 doubles")
        B, keep = x0.shape[0], (int(npts) - 1) // int(stride) + 1
        x = np.empty((keep, n, B))
        u = np.empty((keep, m, B)) if with_inputs else None
        self._ck(self.lib.pdp_rollout(self.h, phys.ctypes.data, x0.ctypes.data, B, int(npts), float(dt), int(stride),
                                      x.ctypes.data, u.ctypes.data if with_inputs else None))
        return x.transpose(2, 0, 1), (u.transpose(2, 0, 1) if with_inputs else None)

    @property
    def nodes_padded(self):
        return int(self.lib.pdp_nodes_padded(self.h))

    @property
    def launch_count(self):
        return int(self.lib.pdp_launch_count(self.h))

    @property
    def last_sweep_ms(self):
        return float(self.lib.pdp_last_sweep_ms(self.h))


class MultiEngine:
    """The Engine interface over n slab handles driven by ONE host thread (include/pyrodp.h, pdp_multi_*): the grid is
    cut over axis 0, part i lives on ``devices[i]`` (default: round-robin over all visible GPUs), halo planes travel
    by peer copies under the interior planes.  No launcher, no torch.distributed: a plain script uses every GPU."""

    def __init__(self, problem: Problem, n_parts=None, devices=None):
        self.lib = _lib.load()
        self.problem = problem
        self.N, self.A, self.n, self.m = problem.N, problem.A, problem.n, problem.m
        ndev = int(self.lib.pdp_device_count())
        if devices is not None:
            n_parts = len(devices)
        elif n_parts is None:
            n_parts = max(ndev, 1)
        dev_arr = (C.c_int32 * n_parts)(*devices) if devices is not None else None
        h = C.c_void_p()
        rc = self.lib.pdp_multi_create(C.byref(problem.c), int(n_parts), dev_arr, C.byref(h))
        if rc != _lib.PDP_OK:
            msg = self.lib.pdp_multi_last_error(None)
            _lib.check(rc) if not msg else _raise(rc, msg.decode())
        self.h = h
        self.n_parts = int(self.lib.pdp_multi_parts(self.h))
        self.devices = [int(self.lib.pdp_multi_part_device(self.h, i)) for i in range(self.n_parts)]
        self.n0 = problem.dims[0]
        self.plane = self.N // self.n0
        self.slab_begin, self.slab_end, self.alloc_begin, self.alloc_end = 0, self.n0, 0, self.n0
        self.slab_nodes = self.N
        self.layouts = []
        for i in range(self.n_parts):
            lay = (C.c_int32 * 8)()
            _lib.check(self.lib.pdp_slab_layout(self.lib.pdp_multi_part(self.h, i), lay))
            self.layouts.append(tuple(int(v) for v in lay))
        self.halo_lo, self.halo_hi = self.layouts[0][4], self.layouts[0][5]
        self.lanes_per_node = self.layouts[0][7]
        self._enqueued = 0

    def _ck(self, rc):
        if rc != _lib.PDP_OK:
            msg = self.lib.pdp_multi_last_error(self.h)
            _raise(rc, msg.decode() if msg else f"pdp error {rc}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pdp_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval_terminal_cost(self):
        self._ck(self.lib.pdp_multi_eval_terminal_cost(self.h))

    def set_J(self, J):
        J = np.ascontiguousarray(J, dtype=np.float64)
        if J.size != self.N:
            raise ValueError("Grid size does not match data")
        self._ck(self.lib.pdp_multi_set_J(self.h, J.ctypes.data))

    def _get(self, which, dtype, out):
        if out is None:
            out = np.empty(self.N, dtype=dtype)
        elif out.size != self.N or out.dtype != dtype or not out.flags.c_contiguous:
            raise ValueError("output buffer does not match the grid size / dtype")
        self._ck(self.lib.pdp_multi_get(self.h, which, out.ctypes.data))
        return out

    def get_J(self, out=None):
        return self._get(0, np.float64, out)

    def get_J_next(self, out=None):
        return self._get(1, np.float64, out)

    def get_pi(self, out=None):
        return self._get(2, np.int64, out)

    def sweep(self, n_sweeps=1):
        stats = np.empty((max(n_sweeps, 0), 3), dtype=np.float64)
        self._ck(self.lib.pdp_multi_sweep(self.h, int(n_sweeps), stats.ctypes.data))
        return stats

    def sweep_nowait(self):
        self._ck(self.lib.pdp_multi_sweep_enqueue(self.h))
        self._enqueued += 1

    def collect_stats(self):
        n = self._enqueued
        out = np.empty((max(n, 1), 3), dtype=np.float64)
        got = C.c_int32(0)
        self._ck(self.lib.pdp_multi_sweep_collect(self.h, out.ctypes.data, n, C.byref(got)))
        self._enqueued = 0
        return out[:got.value]

    def get_input_from_policy(self, k):
        out = np.empty(self.N, dtype=np.float64)
        self._ck(self.lib.pdp_multi_get_input_from_policy(self.h, int(k), out.ctypes.data))
        return out

    def clean_infeasible_set(self, tol, default_action):
        self._ck(self.lib.pdp_multi_clean_infeasible_set(self.h, float(tol), int(default_action)))

    @property
    def kernel_info(self):
        buf = C.create_string_buffer(160)
        _lib.check(self.lib.pdp_kernel_info(self.lib.pdp_multi_part(self.h, 0), buf, 160))
        return buf.value.decode()

    @property
    def launch_count(self):
        return int(self.lib.pdp_multi_launch_count(self.h))


def _raise(rc, msg):
    if rc == _lib.PDP_EINVAL:
        raise ValueError(msg)
    if rc == _lib.PDP_ENOTSUP:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def device_count():
    """Visible CUDA devices (0 without a driver); from the library, so torch is not needed."""
    return int(_lib.load().pdp_device_count())
