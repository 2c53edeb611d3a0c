"""CPU stand-in with the ``Engine`` interface, backed by the C oracle.

Used ONLY to test host-side logic without a GPU (the DynamicProgramming bookkeeping and the
multi-rank slab/exchange logic over gloo).  It is test infrastructure: the product never
constructs it (pyro_b200 has no CPU path).

It mirrors the library's memory layout: the J buffers hold planes [alloc_begin, alloc_end) only
(slab + halo, halo from the library's own ``pdp_compute_halo``).  Every backup is run by the
oracle on a full-size array that is NaN outside those planes, so a halo that is too narrow, or
a halo plane that was not exchanged, poisons the result and fails the test.
"""
import ctypes as C

import numpy as np
import torch

from oracle import c_oracle
from pyro_b200 import _lib


class FakeEngine:
    def __init__(self, problem):
        self.problem = problem
        self.N, self.A, self.n, self.m = problem.N, problem.A, problem.n, problem.m
        c = problem.c
        self.n0 = c.dims[0]
        self.plane = self.N // self.n0
        self.slab_begin, self.slab_end = c.slab_begin, c.slab_end
        lo, hi = C.c_int32(), C.c_int32()
        _lib.check(_lib.load().pdp_compute_halo(C.byref(c), C.byref(lo), C.byref(hi)))
        self.halo_lo, self.halo_hi = lo.value, hi.value
        partial = not (c.slab_begin == 0 and c.slab_end == self.n0)
        if not partial or c.alloc_planes > 0 or c.system_id == _lib.PDP_SYS_LUT:
            self.alloc_begin, self.alloc_end = 0, self.n0
            cap = c.alloc_planes if c.alloc_planes > 0 else self.n0
        else:
            self.alloc_begin = max(0, c.slab_begin - self.halo_lo)
            self.alloc_end = min(self.n0, c.slab_end + self.halo_hi)
            cap = self.alloc_end - self.alloc_begin
        self.nodes_padded = cap * self.plane
        self.lo, self.hi = c.slab_begin * self.plane, c.slab_end * self.plane
        self.slab_nodes = self.hi - self.lo
        self.a0, self.a1 = self.alloc_begin * self.plane, self.alloc_end * self.plane
        self.J = [torch.full((self.nodes_padded,), float("nan"), dtype=torch.float64) for _ in range(2)]
        self.cur = 0
        self.pi = torch.zeros(max(self.slab_nodes, 1), dtype=torch.int64)
        self.stats = torch.zeros(12, dtype=torch.float64)
        self.launch_count = 0
        self.pending = False

    def _held(self, which):
        return self.J[which][:self.a1 - self.a0]

    def _slab(self, which):
        return self.J[which][self.lo - self.a0:self.hi - self.a0]

    # -- Engine interface --
    def eval_terminal_cost(self):
        self._held(self.cur)[:] = torch.from_numpy(c_oracle.terminal(self.problem)[self.a0:self.a1])

    def set_J(self, J):
        if np.size(J) != self.N:
            raise ValueError("Grid size does not match data")
        self._held(self.cur)[:] = torch.from_numpy(np.ascontiguousarray(J, dtype=np.float64)[self.a0:self.a1])

    def get_J(self, out=None):
        return self._slab(self.cur).numpy().copy()

    def get_J_next(self, out=None):
        return self._slab(1 - self.cur).numpy().copy()

    def get_pi(self, out=None):
        return self.pi[:self.slab_nodes].numpy().copy()

    def sweep_async(self):
        Jn = np.full(self.N, np.nan)
        Jn[self.a0:self.a1] = self._held(self.cur).numpy()
        if self.hi > self.lo:
            J, pi = c_oracle.sweep_fused(self.problem, Jn, self.lo, self.hi, n_threads=1)
            assert not np.isnan(J).any(), "backup read a plane outside slab + halo (or a stale halo)"
            self._slab(1 - self.cur)[:] = torch.from_numpy(J)
            self.pi[:self.slab_nodes] = torch.from_numpy(pi)
            d = J - Jn[self.lo:self.hi]
            self.stats[:3] = torch.tensor([J.max(), d.max(), d.min()])
        else:
            self.stats[:3] = torch.tensor([-np.inf, -np.inf, np.inf])
        self.launch_count += 1
        self.pending = True

    def commit_sweep(self):
        assert self.pending
        self.cur = 1 - self.cur
        self.pending = False

    def sweep(self, n_sweeps=1):
        out = np.empty((n_sweeps, 3))
        for k in range(n_sweeps):
            self.sweep_async()
            out[k] = self.stats[:3].numpy()
            self.commit_sweep()
        return out

    def device_buffers(self):
        return ("J_cur", "J_new", "pi", "stats")

    def wrap(self, token, count, typestr):
        return {"J_cur": self.J[self.cur], "J_new": self.J[1 - self.cur], "pi": self.pi, "stats": self.stats}[token]

    def get_input_from_policy(self, k):
        U = np.stack([g.reshape(-1) for g in np.meshgrid(
            *[self.problem.tables[f"u_level{i}"] for i in range(self.m)], indexing="ij")], axis=1)
        return U[self.get_pi(), k]

    def clean_infeasible_set(self, tol, default_action):
        J = self._slab(self.cur)
        bad = J > (self.problem.c.INF - tol)
        J[bad] = self.problem.c.INF
        self.pi[:self.slab_nodes][bad] = default_action

    def close(self):
        pass
