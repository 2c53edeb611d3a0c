"""CPU stand-in with the ``Engine`` interface, backed by the C oracle.

Used ONLY to test host-side logic without a GPU (the DynamicProgramming bookkeeping and the
multi-rank slab/exchange logic over gloo).  It is test infrastructure: the product never
constructs it (pyro_b200 has no CPU path).
"""
import numpy as np
import torch

from oracle import c_oracle


class FakeEngine:
    def __init__(self, problem):
        self.problem = problem
        self.N, self.A, self.n, self.m = problem.N, problem.A, problem.n, problem.m
        c = problem.c
        self.plane = self.N // c.dims[0]
        self.lo, self.hi = c.slab_begin * self.plane, c.slab_end * self.plane
        planes = c.alloc_planes if c.alloc_planes else c.dims[0]
        self.n_pad = planes * self.plane
        self.J = [torch.zeros(self.n_pad, dtype=torch.float64), torch.zeros(self.n_pad, dtype=torch.float64)]
        self.cur = 0
        self.pi = torch.zeros(self.N, dtype=torch.int64)
        self.stats = torch.zeros(3, dtype=torch.float64)
        self.launch_count = 0
        self.pending = False

    # -- Engine interface --
    def eval_terminal_cost(self):
        self.J[self.cur][:self.N] = torch.from_numpy(c_oracle.terminal(self.problem))

    def set_J(self, J):
        if np.size(J) != self.N:
            raise ValueError("Grid size does not match data")
        self.J[self.cur][:self.N] = torch.from_numpy(np.ascontiguousarray(J, dtype=np.float64))

    def get_J(self, out=None):
        return self.J[self.cur][:self.N].numpy().copy()

    def get_J_next(self, out=None):
        return self.J[1 - self.cur][:self.N].numpy().copy()

    def get_pi(self, out=None):
        return self.pi.numpy().copy()

    def sweep_async(self):
        Jn = self.J[self.cur][:self.N].numpy()
        if self.hi > self.lo:
            J, pi = c_oracle.sweep_fused(self.problem, Jn, self.lo, self.hi, n_threads=1)
            self.J[1 - self.cur][self.lo:self.hi] = torch.from_numpy(J)
            self.pi[self.lo:self.hi] = torch.from_numpy(pi)
            d = J - Jn[self.lo:self.hi]
            self.stats[:] = torch.tensor([J.max(), d.max(), d.min()])
        else:
            self.stats[:] = torch.tensor([-np.inf, -np.inf, np.inf])
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
            out[k] = self.stats.numpy()
            self.commit_sweep()
        return out

    def device_buffers(self):
        return ("J_cur", "J_new", "pi", "stats")

    def wrap(self, token, count, typestr):
        return {"J_cur": self.J[self.cur], "J_new": self.J[1 - self.cur], "pi": self.pi, "stats": self.stats}[token]

    def get_input_from_policy(self, k):
        U = np.stack([g.reshape(-1) for g in np.meshgrid(
            *[self.problem.tables[f"u_level{i}"] for i in range(self.m)], indexing="ij")], axis=1)
        return U[self.pi.numpy(), k]

    def clean_infeasible_set(self, tol, default_action):
        J = self.J[self.cur][:self.N]
        bad = J > (self.problem.c.INF - tol)
        J[bad] = self.problem.c.INF
        self.pi[bad] = default_action

    def close(self):
        pass
