"""Tier-1 oracle: NumPy restatement of the reference's grid-DP Bellman sweep.

TEST INFRASTRUCTURE ONLY — never imported by ``pyro_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import this module, and only as the checker / the baseline being timed.

Parity status: PINNED.  The reference has no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the pin is the reference itself: ``oracle/gen_golden.py`` runs the
unmodified reference (tier-0, ``oracle/ref_loader.py``) in the build container and commits its
outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against those
fixtures bit for bit (tables, J and pi) and, when the reference is importable, against live runs.

What is restated, with the reference lines each piece follows:

  levels            np.linspace(lb, ub, dim)                         discretizer.py:134-163
  node/action order C order, last axis fastest                       discretizer.py:167-310
  dynamics          x_next = f(x,u)*dt + x, closed-form f            discretizer.py:363; mechanical.py:222-263;
                                                                      pendulum.py:52-150, 362-493; cartpole.py:335-437;
                                                                      manipulator.py:197-218, 821-992
  validity          strict box tests                                  system.py:198-215
  stage cost table  G = g(x,u)*dt, or INF when invalid               dynamicprogramming.py:517-553; costfunction.py:151-204, 287-334
  interpolation     scipy RegularGridInterpolator, linear, fill 0    discretizer.py:570-587 -> scipy/_rgi.py:375-483, 520-549, 635-642
                    (third-party: scipy, unpinned by the reference ``setup.py:21 scipy>=1.5.2``; the build image has 1.18.1.
                     Published algorithm restated in ``rgi_linear`` below and cross-checked against the installed scipy.)
  backup            Q = G + alpha*J(x_next); J = min, pi = argmin     dynamicprogramming.py:557-570
  statistics        max J, max/min (J - J_next)                       dynamicprogramming.py:247-261

BLAS conventions.  ``np.dot`` on 2-vectors / 2x2 matrices goes through OpenBLAS kernels that use
fused multiply-adds in a fixed order (measured in the build container, OpenBLAS 0.3.30 /
SkylakeX: matvec row = fma(m0, v0, m1*v1); ddot = forward fma chain; 4x4 matvec row =
(p0+p2)+(p1+p3) with separately rounded products).  ``_fma`` reproduces an IEEE fma with a
tiny C helper so the restatement stays bit-identical to the reference.
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------------
# fma helper (C, compiled on first use by oracle/Makefile via __graft_entry__.build())
# ------------------------------------------------------------------------------------------------
_clib = None


def clib():
    global _clib
    if _clib is None:
        so = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", _HERE, "lib"])
        _clib = ctypes.CDLL(so)
        _clib.orc_vfma.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 4
    return _clib


def _fma(a, b, c):
    a, b, c = np.broadcast_arrays(np.asarray(a, float), np.asarray(b, float), np.asarray(c, float))
    a, b, c = (np.ascontiguousarray(v) for v in (a, b, c))
    out = np.empty_like(a)
    clib().orc_vfma(a.size, a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data)
    return out


def _mv2(m0, m1, v0, v1):
    """Row of np.dot(M(2x2), v): fma(m0, v0, m1*v1)."""
    return _fma(m0, v0, np.asarray(m1) * np.asarray(v1))


# ------------------------------------------------------------------------------------------------
# systems: closed-form f(x,u) on arrays, reference expression order
# ------------------------------------------------------------------------------------------------
class SysSpec:
    """Plain parameter record; defaults copied from the reference's constructors."""

    def __init__(self, kind, **kw):
        self.kind = kind
        if kind == "SinglePendulum":            # pendulum.py:52-66; bounds mechanical.py:59-74
            self.n, self.m = 2, 1
            self.par = dict(l1=2.0, lc1=1, m1=1, I1=1, gravity=9.81, d1=0)
            self.u_lb, self.u_ub = np.array([-5.0]), np.array([5.0])
        elif kind == "DoublePendulum":          # pendulum.py:362-378
            self.n, self.m = 4, 2
            self.par = dict(l1=1, l2=1, lc1=1, lc2=1, m1=1, I1=0, m2=1, I2=0, gravity=9.81, d1=0, d2=0)
            self.u_lb, self.u_ub = np.array([-5.0, -5.0]), np.array([5.0, 5.0])
        elif kind == "TwoLinkManipulator":      # manipulator.py:821-837
            self.n, self.m = 4, 2
            self.par = dict(l1=0.5, l2=0.3, lc1=0.2, lc2=0.1, m1=1, I1=0, m2=1, I2=0, gravity=9.81, d1=0.5, d2=0.5)
            self.u_lb, self.u_ub = np.array([-5.0, -5.0]), np.array([5.0, 5.0])
        elif kind == "CartPole":                # cartpole.py:335-359
            self.n, self.m = 4, 1
            self.par = dict(l=3, lcg=0.5, m1=1, m2=0.1, gravity=9.81)
            self.u_lb, self.u_ub = np.array([-10.0]), np.array([10.0])
        else:
            raise ValueError(kind)
        self.x_lb = np.zeros(self.n) - np.pi * 2
        self.x_ub = np.zeros(self.n) + np.pi * 2
        self.xbar = np.zeros(self.n)
        self.ubar = np.zeros(self.m)
        self.par.update(kw)


def _inv2x2_batched(H):
    """np.linalg.inv on a stack of matrices = the same LAPACK gesv call per matrix (mechanical.py:231)."""
    return np.linalg.inv(H)


def f_batch(spec, X, U):
    """dx = f(x, u) for broadcastable state / input arrays X[...,n], U[...,m]."""
    p = spec.par
    if spec.kind == "SinglePendulum":
        q, dq, u = X[..., 0], X[..., 1], U[..., 0]
        Hinv = np.linalg.inv(np.array([[p["m1"] * p["lc1"] ** 2 + p["I1"]]], dtype=float))[0, 0]
        g = p["m1"] * p["gravity"] * p["lc1"] * np.sin(q)
        d = p["d1"] * dq
        rhs = ((1.0 * u - 0.0 * dq) - g) - d          # B u - C dq - g - d, B = [[1]], C = [[0]]
        ddq = Hinv * rhs
        return np.stack(np.broadcast_arrays(dq, ddq), axis=-1)
    if spec.kind in ("DoublePendulum", "TwoLinkManipulator"):
        q0, q1, dq0, dq1 = (X[..., i] for i in range(4))
        u0, u1 = U[..., 0], U[..., 1]
        c2, s2 = np.cos(q1), np.sin(q1)
        s1, s12 = np.sin(q0), np.sin(q0 + q1)
        H = np.zeros(np.shape(c2) + (2, 2))
        H[..., 0, 0] = (p["m1"] * p["lc1"] ** 2 + p["I1"]
                        + p["m2"] * (p["l1"] ** 2 + p["lc2"] ** 2 + 2 * p["l1"] * p["lc2"] * c2) + p["I2"])
        H[..., 1, 0] = p["m2"] * p["lc2"] ** 2 + p["m2"] * p["l1"] * p["lc2"] * c2 + p["I2"]
        H[..., 0, 1] = H[..., 1, 0]
        H[..., 1, 1] = p["m2"] * p["lc2"] ** 2 + p["I2"]
        Hinv = _inv2x2_batched(H)
        h = p["m2"] * p["l1"] * p["lc2"] * s2
        C00, C10, C01 = -h * dq1, h * dq0, -h * (dq0 + dq1)
        cd0 = _mv2(C00, C01, dq0, dq1)
        cd1 = _mv2(C10, 0.0, dq0, dq1)
        g1 = (p["m1"] * p["lc1"] + p["m2"] * p["l1"]) * p["gravity"]
        g2 = p["m2"] * p["lc2"] * p["gravity"]
        G0 = -g1 * s1 - g2 * s12
        G1 = -g2 * s12
        d0 = _mv2(p["d1"], 0.0, dq0, dq1)
        d1 = _mv2(0.0, p["d2"], dq0, dq1)
        bu0 = _mv2(1.0, 0.0, u0, u1)
        bu1 = _mv2(0.0, 1.0, u0, u1)
        r0 = ((bu0 - cd0) - G0) - d0
        r1 = ((bu1 - cd1) - G1) - d1
        ddq0 = _mv2(Hinv[..., 0, 0], Hinv[..., 0, 1], r0, r1)
        ddq1 = _mv2(Hinv[..., 1, 0], Hinv[..., 1, 1], r0, r1)
        return np.stack(np.broadcast_arrays(dq0, dq1, ddq0, ddq1), axis=-1)
    if spec.kind == "CartPole":
        q1, dq0, dq1 = X[..., 1], X[..., 2], X[..., 3]
        u0 = U[..., 0]
        H = np.zeros(np.shape(q1) + (2, 2))
        H[..., 0, 0] = p["m1"] + p["m2"]
        H[..., 1, 0] = p["m2"] * p["lcg"] * np.cos(q1)
        H[..., 0, 1] = H[..., 1, 0]
        H[..., 1, 1] = p["m2"] * p["lcg"] ** 2
        Hinv = _inv2x2_batched(H)
        C01 = -p["m2"] * p["lcg"] * np.sin(q1) * dq1
        cd0 = _mv2(0.0, C01, dq0, dq1)
        cd1 = _mv2(0.0, 0.0, dq0, dq1)
        G1 = p["m2"] * p["gravity"] * p["lcg"] * np.sin(q1)
        bu0, bu1 = 1.0 * u0, 0.0 * u0               # np.dot(B(2x1), u(1,))
        r0 = ((bu0 - cd0) - 0.0) - 0.0
        r1 = ((bu1 - cd1) - G1) - 0.0
        ddq0 = _mv2(Hinv[..., 0, 0], Hinv[..., 0, 1], r0, r1)
        ddq1 = _mv2(Hinv[..., 1, 0], Hinv[..., 1, 1], r0, r1)
        return np.stack(np.broadcast_arrays(dq0, dq1, ddq0, ddq1), axis=-1)
    raise ValueError(spec.kind)


# ------------------------------------------------------------------------------------------------
# cost functions on arrays
# ------------------------------------------------------------------------------------------------
class QuadCost:
    """QuadraticCostFunction (costfunction.py:100-204)."""

    def __init__(self, n, m, xbar=None, ubar=None):
        self.n, self.m = n, m
        self.xbar = np.zeros(n) if xbar is None else np.asarray(xbar, float)
        self.ubar = np.zeros(m) if ubar is None else np.asarray(ubar, float)
        self.Q, self.R, self.S = np.diag(np.ones(n)), np.diag(np.ones(m)), np.diag(np.zeros(n))
        self.INF, self.EPS, self.ontarget_check = 1e3, 1e-3, True

    def _quad(self, W, d):
        """np.dot(d.T, np.dot(W, d)) on arrays d[..., k] with the measured BLAS association."""
        k = d.shape[-1]
        if k == 1:
            return d[..., 0] * (W[0, 0] * d[..., 0])
        if k == 2:
            w0 = _mv2(W[0, 0], W[0, 1], d[..., 0], d[..., 1])
            w1 = _mv2(W[1, 0], W[1, 1], d[..., 0], d[..., 1])
            return _fma(d[..., 1], w1, d[..., 0] * w0)
        if k == 4:
            w = []
            for i in range(4):
                pr = [W[i, j] * d[..., j] for j in range(4)]
                w.append((pr[0] + pr[2]) + (pr[1] + pr[3]))
            acc = d[..., 0] * w[0]
            for i in range(1, 4):
                acc = _fma(d[..., i], w[i], acc)
            return acc
        raise NotImplementedError

    def _norm(self, d):
        acc = d[..., 0] * d[..., 0]
        for i in range(1, d.shape[-1]):
            acc = _fma(d[..., i], d[..., i], acc)
        return np.sqrt(acc)

    def g(self, X, U):
        dx, du = X - self.xbar, U - self.ubar
        dJ = self._quad(self.Q, dx) + self._quad(self.R, du)
        if self.ontarget_check:
            dJ = np.where(self._norm(dx) < self.EPS, 0.0, dJ)
        return dJ

    def h(self, X):
        dx = X - self.xbar
        J = self._quad(self.S, dx)
        if self.ontarget_check:
            J = np.where(self._norm(dx) < self.EPS, 0.0, J)
        return J


class TimeCost(QuadCost):
    """TimeCostFunction (costfunction.py:287-334): g = 1 (0 on target), h = 0."""

    def __init__(self, xbar):
        xbar = np.asarray(xbar, float)
        super().__init__(xbar.size, 1, xbar)

    def g(self, X, U):
        dJ = np.ones(np.broadcast_shapes(X.shape[:-1], U.shape[:-1]))
        if self.ontarget_check:
            dJ = np.where(self._norm(X - self.xbar) < self.EPS, 0.0, dJ)
        return dJ

    def h(self, X):
        return np.zeros(X.shape[:-1])


class ReachCost(QuadCost):
    """Reachability (costfunction.py:421-481) with the system's box isavalidstate and the default norm test: g = 0 on a node
    inside the box (grid nodes always are), h = 0 if ||x - xbar|| < EPS else INF.  INF = 1e4, EPS = 0.2 by default."""

    def __init__(self, xbar):
        xbar = np.asarray(xbar, float)
        super().__init__(xbar.size, 1, xbar)
        self.INF, self.EPS = 1e4, 0.2

    def g(self, X, U):
        return np.zeros(np.broadcast_shapes(X.shape[:-1], U.shape[:-1]))

    def h(self, X):
        return np.where(self._norm(X - self.xbar) < self.EPS, 0.0, self.INF)


# ------------------------------------------------------------------------------------------------
# RegularGridInterpolator(method='linear', bounds_error=False, fill_value=0) restated
# ------------------------------------------------------------------------------------------------
def find_indices(level, x):
    """scipy find_indices: i = clip(searchsorted(level, x, 'right') - 1, 0, len-2); y = (x-l[i])/(l[i+1]-l[i])."""
    i = np.clip(np.searchsorted(level, x, side="right") - 1, 0, level.size - 2)
    y = (x - level[i]) / (level[i + 1] - level[i])
    return i, y


def rgi_linear(levels, values, xi):
    """values: n-D grid; xi: (..., n).  Returns interpolated values, 0 where out of bounds."""
    n = len(levels)
    shape = xi.shape[:-1]
    xi = xi.reshape(-1, n)
    oob = np.zeros(xi.shape[0], dtype=bool)
    idx, y = [], []
    for d in range(n):
        oob |= xi[:, d] < levels[d][0]
        oob |= xi[:, d] > levels[d][-1]
        i, yy = find_indices(levels[d], xi[:, d])
        idx.append(i)
        y.append(yy)
    if n == 2:  # evaluate_linear_2d: value-first products, terms added left to right
        i0, i1, y0, y1 = idx[0], idx[1], y[0], y[1]
        out = values[i0, i1] * (1 - y0) * (1 - y1)
        out = out + values[i0, i1 + 1] * (1 - y0) * y1
        out = out + values[i0 + 1, i1] * y0 * (1 - y1)
        out = out + values[i0 + 1, i1 + 1] * y0 * y1
    else:       # _evaluate_linear: weight-first, itertools.product corner order
        out = np.zeros(xi.shape[0])
        for corner in itertools.product((0, 1), repeat=n):
            w = np.ones(xi.shape[0])
            for d, bit in enumerate(corner):
                w = w * (y[d] if bit else (1 - y[d]))
            out = out + values[tuple(idx[d] + corner[d] for d in range(n))] * w
    out[oob] = 0.0
    return out.reshape(shape)


# ------------------------------------------------------------------------------------------------
# the grid + tables + sweep
# ------------------------------------------------------------------------------------------------
class GridOracle:
    def __init__(self, spec, x_grid_dim, u_grid_dim, dt=0.05):
        self.spec, self.dt = spec, dt
        self.dims, self.udims = tuple(int(d) for d in x_grid_dim), tuple(int(d) for d in u_grid_dim)
        self.x_level = [np.linspace(spec.x_lb[i], spec.x_ub[i], self.dims[i]) for i in range(spec.n)]
        self.u_level = [np.linspace(spec.u_lb[i], spec.u_ub[i], self.udims[i]) for i in range(spec.m)]
        self.N, self.A = int(np.prod(self.dims, dtype=np.int64)), int(np.prod(self.udims))
        self.U = np.stack([g.reshape(-1) for g in np.meshgrid(*self.u_level, indexing="ij")], axis=1)

    def states(self, lo=0, hi=None):
        """state_from_node_id[lo:hi] without materialising the whole table."""
        hi = self.N if hi is None else hi
        ids = np.arange(lo, hi)
        idx = np.unravel_index(ids, self.dims)
        return np.stack([self.x_level[d][idx[d]] for d in range(self.spec.n)], axis=1)

    def tables(self, cost, lo=0, hi=None):
        """x_next (K,A,n), x_ok, a_ok (K,A) bool, G (K,A) for nodes lo:hi."""
        X = self.states(lo, hi)[:, None, :]
        U = self.U[None, :, :]
        x_next = f_batch(self.spec, X, U) * self.dt + X
        s = self.spec
        x_ok = ~np.any((x_next < s.x_lb) | (x_next > s.x_ub), axis=-1)
        a_ok = np.broadcast_to(~np.any((U < s.u_lb) | (U > s.u_ub), axis=-1), x_ok.shape)
        G = np.where(a_ok & x_ok, cost.g(X, U) * self.dt, cost.INF)
        return x_next, x_ok, a_ok, G

    def terminal(self, cost):
        return np.asarray(cost.h(self.states()), dtype=float)

    def sweep(self, J_next, cost, alpha=1.0, chunk=1 << 15, use_scipy=False, lo=0, hi=None):
        """One Bellman backup for nodes lo:hi, LUT formulation chunked over nodes."""
        hi = self.N if hi is None else hi
        grid = J_next.reshape(self.dims)
        interp = None
        if use_scipy:
            from scipy.interpolate import RegularGridInterpolator
            interp = RegularGridInterpolator(tuple(self.x_level), grid, "linear", False, 0)
        J = np.empty(hi - lo)
        pi = np.empty(hi - lo, dtype=np.int64)
        step = max(1, chunk // self.A)
        for a in range(lo, hi, step):
            b = min(hi, a + step)
            x_next, _, _, G = self.tables(cost, a, b)
            Jx = interp(x_next) if use_scipy else rgi_linear(self.x_level, grid, x_next)
            Q = G + alpha * Jx
            J[a - lo:b - lo] = Q.min(axis=1)
            pi[a - lo:b - lo] = Q.argmin(axis=1)
        return J, pi

    def run(self, cost, n_sweeps, alpha=1.0, J0=None, **kw):
        J = self.terminal(cost) if J0 is None else np.array(J0, float)
        pi = np.zeros(self.N, dtype=np.int64)
        stats = []
        for _ in range(n_sweeps):
            Jn = J
            J, pi = self.sweep(Jn, cost, alpha, **kw)
            d = J - Jn
            stats.append((J.max(), d.max(), d.min()))
        return J, pi, np.array(stats)


def lut_sweep(x_level, dims, J_next, x_next, G, alpha=1.0, use_scipy=True):
    """The reference's three NumPy lines on given tables (dynamicprogramming.py:564-570)."""
    grid = J_next.reshape(dims)
    if use_scipy:
        from scipy.interpolate import RegularGridInterpolator
        Jx = RegularGridInterpolator(tuple(x_level), grid, "linear", False, 0)(x_next)
    else:
        Jx = rgi_linear(x_level, grid, x_next)
    Q = G + alpha * Jx
    return Q.min(axis=1), Q.argmin(axis=1)


def spline_sweep(x_level, dims, J_next, x_next, G, alpha=1.0):
    """One backup of the reference's DynamicProgramming2DRectBivariateSpline on given tables
    (dynamicprogramming.py:582-614; interpolant: discretizer.py:591-612 = scipy RectBivariateSpline kx = ky = 3, s = 0 —
    third-party FITPACK, executed, not restated).  Returns (J, pi, gap between the best and second-best Q per node)."""
    from scipy.interpolate import RectBivariateSpline
    interp = RectBivariateSpline(x_level[0], x_level[1], J_next.reshape(dims), bbox=[None, None, None, None], kx=3, ky=3)
    Jx = interp(x_next[:, :, 0].flatten(), x_next[:, :, 1].flatten(), grid=False).reshape(G.shape)
    Q = G + alpha * Jx
    Qs = np.sort(Q, axis=1)
    return Q.min(axis=1), Q.argmin(axis=1), Qs[:, 1] - Qs[:, 0]
