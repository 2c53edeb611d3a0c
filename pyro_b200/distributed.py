"""Multi-GPU value iteration: slabs over the outermost grid axis, one process per GPU.

SURVEY.md section 8e.  Every node's backup reads J_next only (Jacobi style,
pyro/planning/dynamicprogramming.py:181-191), so the node set shards freely and the one
exchange per sweep is the new J near the slab boundaries.

Two backends drive the exchange: "native" (default on GPUs) — the C library owns an NCCL
communicator (``pdp_comm_init``) and runs sweep + exchange + statistics all-reduce inside one call;
"torch" — this module moves the same planes through ``torch.distributed`` on zero-copy views of
the engine's buffers (any backend, e.g. gloo in the CPU tests).

Halo mode (the normal case).  Rank r owns axis-0 planes [r*N0//W, (r+1)*N0//W) and holds only
those planes plus the halo its backups can read: axis 0 is a position, x_next[0] = dq0*dt + q0,
so a node reads at most ``halo_lo`` planes below and ``halo_hi`` planes above its own — computed
exactly from the levels by the library (``pdp_slab_layout``).  Per sweep the boundary planes are
computed first, their exchange with ranks r-1 / r+1 (grouped NCCL send/recv over NVLink, on a
side stream) overlaps the interior planes, and the convergence statistics are one small
all-reduce for the whole batch of sweeps.  Memory per GPU is O(N/W + halo): the 201^4 grid of
BASELINE config 5 needs 2 x 1.6 GB + halo per rank instead of 2 x 13 GB.

All-gather mode (fallback).  When the halo is wider than a neighbour's slab (tiny grids on many
ranks) or unknown (LUT mode: an arbitrary x_next_table), every rank keeps the full J, padded to
W*ceil(N0/W) planes, and one in-place ``all_gather_into_tensor`` completes the new J.

``torch.distributed`` is plumbing only: the tensors are zero-copy views of the engine's own
device buffers.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_sharded():
    try:
        dist = _dist()
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def slab_of(rank, world, n_planes):
    """All-gather mode partition: (begin, end, planes_per_rank); trailing ranks may be short or empty."""
    per = -(-n_planes // world)
    begin = min(rank * per, n_planes)
    end = min(begin + per, n_planes)
    return begin, end, per


def balanced_slab(rank, world, n_planes):
    """Halo mode partition: thickness differs by at most one plane between ranks."""
    return rank * n_planes // world, (rank + 1) * n_planes // world


def rebalanced_bounds(bounds, times, min_thick):
    """New slab boundaries from the measured compute time of each slab (work per plane taken as constant inside a slab):
    equal cumulative work per rank, every slab at least ``min_thick`` planes.  ``bounds``: W+1 plane indices."""
    W = len(bounds) - 1
    n0 = bounds[-1]
    dens = np.zeros(n0)
    for r in range(W):
        if bounds[r + 1] > bounds[r]:
            dens[bounds[r]:bounds[r + 1]] = max(float(times[r]), 1e-9) / (bounds[r + 1] - bounds[r])
    cum = np.concatenate([[0.0], np.cumsum(dens)])
    new = [0]
    for r in range(1, W):
        new.append(int(np.searchsorted(cum, cum[-1] * r / W, side="left")))
    new.append(n0)
    for r in range(1, W):                   # forward / backward passes: minimum thickness
        new[r] = max(new[r], new[r - 1] + min_thick)
    for r in range(W - 1, 0, -1):
        new[r] = min(new[r], new[r + 1] - min_thick)
    if any(new[r + 1] - new[r] < min_thick for r in range(W)):
        return list(bounds)
    return new


class _DevView:
    """Expose a raw device pointer through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _device_tensor(ptr, count, typestr):
    import torch
    return torch.as_tensor(_DevView(ptr, count, typestr), device="cuda")


class ShardedEngine:
    """Same interface as ``engine.Engine`` (sweep / get_J / get_pi / ...), sharded over ranks."""

    def __init__(self, grid_sys, cf, alpha=1.0, interpol_method="linear", engine_factory=None, group=None,
                 mode=None, overlap=True, backend=None, halo=None, bounds=None):
        import torch
        from . import problem as _problem
        dist = _dist()
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._peer = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
        n0 = int(grid_sys.x_grid_dim[0])
        self.n0 = n0
        self.cpu = engine_factory is not None  # CPU stand-in engines (gloo tests of this host logic)

        def make(slab, alloc_planes):
            P = _problem.extract(grid_sys, cf, alpha, interpol_method, slab=slab, alloc_planes=alloc_planes)
            if engine_factory is None:
                from .engine import Engine
                eng = Engine(P)
                eng.set_stream(torch.cuda.current_stream().cuda_stream)
            else:
                eng = engine_factory(P)
            return P, eng

        # halo mode if every rank's halo lies inside its direct neighbours' slabs
        self.mode = None
        min_thick = n0 // self.world
        self.bounds = None
        if mode in (None, "halo") and min_thick >= 1:
            self.bounds = list(bounds) if bounds is not None else [balanced_slab(r, self.world, n0)[0] for r in range(self.world)] + [n0]
            self.begin, self.end = self.bounds[self.rank], self.bounds[self.rank + 1]
            self.problem, self.eng = make((self.begin, self.end), 0)
            fused = self.problem.system_id != 0
            thinnest = min(self.bounds[r + 1] - self.bounds[r] for r in range(self.world))
            if fused and self.eng.halo_lo <= thinnest and self.eng.halo_hi <= thinnest:
                self.mode = "halo"
            else:
                if mode == "halo":
                    raise ValueError("halo exchange impossible: halo wider than a slab, or LUT mode")
                self.eng.close()
        if self.mode is None:
            self.mode = "allgather"
            self.begin, self.end, self.per = slab_of(self.rank, self.world, n0)
            self.problem, self.eng = make((self.begin, self.end), self.per * self.world)
        self.N, self.A, self.n, self.m = self.problem.N, self.problem.A, self.problem.n, self.problem.m
        self.plane = self.N // n0
        self.slab_nodes = (self.end - self.begin) * self.plane
        self.alloc_begin, self.alloc_end = self.eng.alloc_begin, self.eng.alloc_end
        self.halo_lo, self.halo_hi = self.eng.halo_lo, self.eng.halo_hi
        self.n_alloc = self.eng.nodes_padded
        self.overlap = bool(overlap) and self.mode == "halo" and not self.cpu
        self.exchanges = 0
        # backend "native": the library drives NCCL itself (pdp_comm_init) — sweep, exchange and the
        # statistics all-reduce are one C call per batch.  backend "torch": this module drives the
        # exchange through torch.distributed on views of the engine's buffers (any transport; used by
        # the gloo tests of the slab / halo logic with CPU stand-in engines).
        self.backend = backend or ("torch" if self.cpu else "native")
        self.halo = None
        if self.backend == "native":
            ident = [self.eng.nccl_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ident, src=self._peer(0), group=group)
            self.eng.comm_init(self.rank, self.world, ident[0], self.mode, self.overlap)
            # halo planes by grouped NCCL send/recv (default), or by peer-memory stores over NVLink from a device kernel
            # (halo="peer" / PYRODP_HALO=peer).  Measured on 2 B200s at cfg 2 (profiles/r01_halo_ab.txt): NCCL 0.350 ms/step on
            # both ranks, run after run; peer stores 0.342-0.344 on the rank that never waits but 0.344-0.377 on the other
            # (the device-side wait sits at the head of the step on the main stream) — so NCCL stays the default.
            import os
            self.halo = halo or os.environ.get("PYRODP_HALO", "nccl")
            if self.mode == "halo" and self.halo == "peer" and self.world > 1:
                exports = [None] * self.world
                dist.all_gather_object(exports, self.eng.peer_export(), group=group)
                self.eng.peer_attach(exports[self.rank - 1] if self.rank > 0 else None,
                                     exports[self.rank + 1] if self.rank < self.world - 1 else None)
                dist.barrier(group=group)   # every rank has its neighbours' buffers mapped before the first push
            else:
                self.halo = "nccl" if self.mode == "halo" else "allgather"
        elif self.overlap:
            self.comm_stream = torch.cuda.Stream()
            self.ev_boundary = torch.cuda.Event()
            self.ev_comm = torch.cuda.Event()

    # ---- work-balanced slabs -----------------------------------------------------------------------------
    @classmethod
    def balanced(cls, grid_sys, cf, alpha=1.0, interpol_method="linear", group=None, min_gain=0.02, **kw):
        """Halo-mode engine whose slab boundaries equalise the MEASURED compute time per rank: nodes whose position rows
        leave the box cost nothing and are not spread evenly over axis 0 (cfg5 on 8 GPUs: slowest rank 10 % above the
        mean with equal slabs).  One calibration sweep on equal slabs (from h(x); timing does not depend on J), then the
        engine is rebuilt on the new boundaries if that promises more than ``min_gain``."""
        eng = cls(grid_sys, cf, alpha, interpol_method, group=group, **kw)
        if eng.mode != "halo" or eng.world == 1 or eng.cpu or eng.backend != "native":
            return eng
        torch, dist = eng.torch, eng.dist
        eng.eval_terminal_cost()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.eng.sweep_async(); eng.eng.commit_sweep()          # warm-up (first launch of the kernel)
        torch.cuda.synchronize()
        a.record(); eng.eng.sweep_async(); b.record(); eng.eng.commit_sweep()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
        times = [torch.zeros_like(t) for _ in range(eng.world)]
        dist.all_gather(times, t, group=group)
        times = [float(x.item()) for x in times]
        min_thick = max(eng.halo_lo, eng.halo_hi, 1)
        new = rebalanced_bounds(eng.bounds, times, min_thick)
        gain = 1.0 - (sum(times) / eng.world) / max(times)
        if new == eng.bounds or gain < min_gain:
            eng.calibration = {"times_ms": times, "bounds": eng.bounds, "rebalanced": False}
            return eng
        eng.close()
        torch.cuda.empty_cache()
        out = cls(grid_sys, cf, alpha, interpol_method, group=group, bounds=new, **kw)
        out.calibration = {"times_ms": times, "equal_bounds": eng.bounds, "bounds": new, "rebalanced": True}
        return out

    # ---- views of the engine's buffers ------------------------------------------------------------
    def _wrap(self, token, count, typestr):
        if self.cpu:
            return self.eng.wrap(token, count, typestr)
        return _device_tensor(token, count, typestr)

    def _buffers(self):
        jc, jn, pi, st = self.eng.device_buffers()
        return (self._wrap(jc, self.n_alloc, "<f8"), self._wrap(jn, self.n_alloc, "<f8"),
                self._wrap(pi, max(self.slab_nodes, 1), "<i8"), self._wrap(st, 12, "<f8"))

    def _planes(self, buf, p0, p1):
        """View of planes [p0,p1) of a J buffer whose element 0 is plane alloc_begin."""
        return buf[(p0 - self.alloc_begin) * self.plane:(p1 - self.alloc_begin) * self.plane]

    # ---- state -------------------------------------------------------------------------------------
    def eval_terminal_cost(self):
        self.eng.eval_terminal_cost()  # evaluated on slab + halo: no exchange needed

    def set_J(self, J):
        self.eng.set_J(J)              # full (N,) host array; the engine keeps the planes it holds

    # ---- halo exchange -----------------------------------------------------------------------------
    def _halo_ops(self, buf):
        dist = self.dist
        ops = []
        if self.rank > 0:
            peer = self._peer(self.rank - 1)
            ops.append(dist.P2POp(dist.isend, self._planes(buf, self.begin, self.begin + self.halo_hi), peer, self.group))
            ops.append(dist.P2POp(dist.irecv, self._planes(buf, self.begin - self.halo_lo, self.begin), peer, self.group))
        if self.rank < self.world - 1:
            peer = self._peer(self.rank + 1)
            ops.append(dist.P2POp(dist.isend, self._planes(buf, self.end - self.halo_lo, self.end), peer, self.group))
            ops.append(dist.P2POp(dist.irecv, self._planes(buf, self.end, self.end + self.halo_hi), peer, self.group))
        return ops

    def _exchange(self, buf):
        """Complete `buf` (a J buffer whose slab planes are fresh) on this rank."""
        self.exchanges += 1
        if self.mode == "allgather":
            cnt = self.per * self.plane
            off = self.rank * cnt
            # in-place all-gather: rank r's chunk already sits at offset r*cnt of the output
            self.dist.all_gather_into_tensor(buf, buf[off:off + cnt], group=self.group)
            return
        ops = self._halo_ops(buf)
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    # ---- hot path ------------------------------------------------------------------------------------
    def _one_sweep(self):
        """One sweep of the slab + exchange; returns the device tensor [j_max, d_max, -d_min] of the slab."""
        torch = self.torch
        b, e, lo, hi = self.begin, self.end, self.halo_lo, self.halo_hi
        if self.overlap and (b + hi) < (e - lo):
            # boundary planes first; their exchange runs on the side stream under the interior planes
            cur = torch.cuda.current_stream()
            self.eng.sweep_planes_async(b, b + hi, 1)
            self.eng.sweep_planes_async(e - lo, e, 2)
            self.ev_boundary.record(cur)
            self.eng.sweep_planes_async(b + hi, e - lo, 0)
            _, j_new, _, st = self._buffers()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(self.ev_boundary)
                self._exchange(j_new)
                self.ev_comm.record(self.comm_stream)
            cur.wait_event(self.ev_comm)
            used = st.view(4, 3)[:3]
        else:
            self.eng.sweep_async()
            _, j_new, _, st = self._buffers()
            self._exchange(j_new)
            used = st.view(4, 3)[:1]
        self.eng.commit_sweep()
        return torch.stack([used[:, 0].max(), used[:, 1].max(), -(used[:, 2].min())])

    def sweep_host(self, J_next, J_out=None, pi_out=None):
        """One sweep with host arrays: J_next is the full (N,) array, J_out / pi_out receive THIS RANK'S slab.
        The rank uploads the planes it holds (slab + halo) and needs no exchange for this one sweep; the
        statistics returned are those of the slab."""
        return self.eng.sweep_host(J_next, J_out, pi_out)

    def sweep_nowait(self):
        """Enqueue one sweep + exchange without any host synchronisation; pair with collect_stats()."""
        if self.backend == "native":
            return self.eng.sweep_nowait()
        self._pending_stats = getattr(self, "_pending_stats", [])
        self._pending_stats.append(self._one_sweep())

    def collect_stats(self):
        if self.backend == "native":
            return self.eng.collect_stats()
        stats, self._pending_stats = getattr(self, "_pending_stats", []), []
        return self._reduce_stats(stats)

    def sweep(self, n_sweeps=1):
        if self.backend == "native":
            return self.eng.sweep(n_sweeps)
        return self._reduce_stats([self._one_sweep() for _ in range(n_sweeps)])

    def _reduce_stats(self, all_stats):
        torch, dist = self.torch, self.dist
        if not all_stats:
            return np.empty((0, 3))
        red = torch.stack(all_stats)
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=self.group)
        out = red.cpu().numpy().astype(np.float64)
        out[:, 2] = -out[:, 2]
        return out

    # ---- results (gathered onto every rank) ----------------------------------------------------------
    def _gather_slabs(self, mine, dtype):
        """All-gather variable-size slabs (padded to the thickest) and return the full (N,) host array."""
        torch, dist = self.torch, self.dist
        thickest = max(self.bounds[r + 1] - self.bounds[r] for r in range(self.world)) if self.mode == "halo" else -(-self.n0 // self.world)
        cnt = thickest * self.plane
        send = torch.zeros(cnt, dtype=dtype, device=mine.device)
        send[:self.slab_nodes] = mine[:self.slab_nodes]
        full = torch.empty(cnt * self.world, dtype=dtype, device=mine.device)
        dist.all_gather_into_tensor(full, send, group=self.group)
        full = full.cpu().numpy().reshape(self.world, cnt)
        out = np.empty(self.N, dtype=full.dtype)
        for r in range(self.world):
            if self.mode == "halo":
                b, e = self.bounds[r], self.bounds[r + 1]
            else:
                b, e, _ = slab_of(r, self.world, self.n0)
            out[b * self.plane:e * self.plane] = full[r, :(e - b) * self.plane]
        return out

    def _ret(self, res, out):
        if out is not None:
            out[...] = res
            return out
        return res

    def get_J(self, out=None):
        jc, _, _, _ = self._buffers()
        return self._ret(self._gather_slabs(self._planes(jc, self.begin, self.end), self.torch.float64), out)

    def get_J_next(self, out=None):
        _, jn, _, _ = self._buffers()
        return self._ret(self._gather_slabs(self._planes(jn, self.begin, self.end), self.torch.float64), out)

    def get_pi(self, out=None):
        _, _, pi, _ = self._buffers()
        return self._ret(self._gather_slabs(pi, self.torch.int64), out)

    def get_input_from_policy(self, k):
        u_level = [self.problem.tables[f"u_level{i}"] for i in range(self.m)]
        U = np.stack([g.reshape(-1) for g in np.meshgrid(*u_level, indexing="ij")], axis=1)
        return U[self.get_pi(), k]

    def clean_infeasible_set(self, tol, default_action):
        self.eng.clean_infeasible_set(tol, default_action)  # slab only ...
        if self.backend == "native":                         # ... then refresh the neighbours' halos
            return self.eng.exchange_current()
        jc, _, _, _ = self._buffers()
        self._exchange(jc)

    @property
    def launch_count(self):
        return self.eng.launch_count

    def close(self):
        if self.halo == "peer":
            self.dist.barrier(group=self.group)   # no rank frees buffers a neighbour has mapped and may still be storing into
        self.eng.close()
