"""Multi-GPU value iteration: slabs over the outermost grid axis, one process per GPU.

SURVEY.md section 8e.  Every node's backup reads J_next only (Jacobi style,
dynamicprogramming.py:181-191), so the node set shards freely; the one exchange per sweep is
the new J.  Rank r owns axis-0 planes [r*P, min((r+1)*P, N0)), P = ceil(N0 / W); it keeps a
full-size (padded to W*P planes) copy of J_next, computes its slab of J_new in place inside
the full-size "new" buffer, and one in-place ``all_gather_into_tensor`` (NCCL over NVLink)
completes that buffer on every rank.  The convergence statistics are one 3-double all-reduce.

``torch.distributed`` is plumbing only: the tensors are zero-copy views of the engine's own
device buffers and the collective runs on the same stream as the sweep kernel.
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
    """(begin, end, planes_per_rank) of rank's slab; trailing ranks may be short or empty."""
    per = -(-n_planes // world)
    begin = min(rank * per, n_planes)
    end = min(begin + per, n_planes)
    return begin, end, per


class _DevView:
    """Expose a raw device pointer through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _device_tensor(ptr, count, typestr):
    import torch
    return torch.as_tensor(_DevView(ptr, count, typestr), device="cuda")


class ShardedEngine:
    """Same interface as ``engine.Engine`` (sweep / get_J / get_pi / ...), sharded over ranks."""

    def __init__(self, grid_sys, cf, alpha=1.0, interpol_method="linear", engine_factory=None, group=None):
        import torch
        from . import problem as _problem
        dist = _dist()
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        n0 = int(grid_sys.x_grid_dim[0])
        self.begin, self.end, self.per = slab_of(self.rank, self.world, n0)
        self.problem = _problem.extract(grid_sys, cf, alpha, interpol_method, slab=(self.begin, self.end),
                                        alloc_planes=self.per * self.world)
        self.N, self.A, self.n, self.m = self.problem.N, self.problem.A, self.problem.n, self.problem.m
        self.plane = self.N // n0
        if engine_factory is None:
            from .engine import Engine
            self.eng = Engine(self.problem)
            self.eng.set_stream(torch.cuda.current_stream().cuda_stream)
            self._wrap = lambda ptr, count, ts: _device_tensor(ptr, count, ts)
        else:  # CPU stand-in for the gloo tests of this host logic
            self.eng = engine_factory(self.problem)
            self._wrap = self.eng.wrap
        self.n_pad = self.per * self.world * self.plane

    # ---- state (replicated J: every rank uploads / evaluates the full array) ----
    def eval_terminal_cost(self):
        self.eng.eval_terminal_cost()

    def set_J(self, J):
        self.eng.set_J(J)

    def get_J(self, out=None):
        return self.eng.get_J(out)

    def get_J_next(self, out=None):
        return self.eng.get_J_next(out)

    def _buffers(self):
        jc, jn, pi, st = self.eng.device_buffers()
        return (self._wrap(jc, self.n_pad, "<f8"), self._wrap(jn, self.n_pad, "<f8"),
                self._wrap(pi, self.N, "<i8"), self._wrap(st, 3, "<f8"))

    # ---- hot path ----
    def sweep(self, n_sweeps=1):
        torch, dist = self.torch, self.dist
        lo, cnt = self.begin * self.plane, self.per * self.plane
        off = self.rank * cnt
        all_stats = []
        for _ in range(n_sweeps):
            self.eng.sweep_async()
            _, j_new, _, st = self._buffers()
            # in-place all-gather: rank r's chunk already sits at offset r*cnt of the output
            dist.all_gather_into_tensor(j_new, j_new[off:off + cnt], group=self.group)
            self.eng.commit_sweep()
            all_stats.append(torch.stack([st[0], st[1], -st[2]]))
        if not all_stats:
            return np.empty((0, 3))
        red = torch.stack(all_stats)
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=self.group)
        out = red.cpu().numpy().astype(np.float64)
        out[:, 2] = -out[:, 2]
        assert lo == off or self.begin == self.end
        return out

    def get_pi(self, out=None):
        """Gather the policy slabs (int64) onto every rank."""
        torch, dist = self.torch, self.dist
        _, _, pi, _ = self._buffers()
        cnt = self.per * self.plane
        full = torch.zeros(cnt * self.world, dtype=torch.int64, device=pi.device)
        mine = torch.zeros(cnt, dtype=torch.int64, device=pi.device)
        n_mine = (self.end - self.begin) * self.plane
        if n_mine:
            mine[:n_mine] = pi[self.begin * self.plane:self.end * self.plane]
        dist.all_gather_into_tensor(full, mine, group=self.group)
        res = full[:self.N].cpu().numpy()
        if out is not None:
            out[...] = res
            return out
        return res

    def get_input_from_policy(self, k):
        return self.problem.tables_u()[self.get_pi(), k]

    def clean_infeasible_set(self, tol, default_action):
        raise NotImplementedError("clean_infeasible_set on a sharded engine: gather with get_J/get_pi first")

    @property
    def launch_count(self):
        return self.eng.launch_count

    def close(self):
        self.eng.close()
