"""Cost-function parameter holders read by the sweep (pyro/analysis/costfunction.py).

Mirrors attribute names / defaults of
  CostFunction           costfunction.py:19-33   (INF = 1e3, EPS = 1e-3)
  QuadraticCostFunction  costfunction.py:100-204 (Q = I, R = I, S = 0, ontarget_check)
  TimeCostFunction       costfunction.py:287-334
  QuadraticCostFunctionWithDomainCheck  costfunction.py:339-415
  Reachability           costfunction.py:421-481 (INF = 1e4, EPS = 0.2)
so real pyro cost objects and these are interchangeable for ``problem.extract``.
"""
import numpy as np


class CostFunction:
    def __init__(self):
        self.INF = 1e3
        self.EPS = 1e-3

    def h(self, x, t=0):
        raise NotImplementedError

    def g(self, x, u, t=0):
        raise NotImplementedError


class QuadraticCostFunction(CostFunction):
    """g = dx'Q dx + du'R du, h = dx'S dx, both zeroed when ||dx|| < EPS."""

    def __init__(self, n, m):
        CostFunction.__init__(self)     # (not super(): QuadraticCostFunctionWithDomainCheck borrows this constructor)
        self.n, self.m = n, m
        self.xbar = np.zeros(n)
        self.ubar = np.zeros(m)
        self.Q = np.diag(np.ones(n))
        self.R = np.diag(np.ones(m))
        self.S = np.diag(np.zeros(n))
        self.ontarget_check = True

    @classmethod
    def from_sys(cls, sys):
        inst = cls(sys.n, sys.m)
        inst.xbar = sys.xbar
        inst.ubar = sys.ubar
        return inst

    def h(self, x, t=0):
        dx = x - self.xbar
        J_f = np.dot(dx.T, np.dot(self.S, dx))
        if self.ontarget_check and np.linalg.norm(dx) < self.EPS:
            J_f = 0
        return J_f

    def g(self, x, u, t=0):
        dx = x - self.xbar
        du = u - self.ubar
        dJ = np.dot(dx.T, np.dot(self.Q, dx)) + np.dot(du.T, np.dot(self.R, du))
        if self.ontarget_check and np.linalg.norm(dx) < self.EPS:
            dJ = 0
        return dJ


class TimeCostFunction(CostFunction):
    """g = 1 (0 on target), h = 0."""

    def __init__(self, xbar):
        super().__init__()
        self.xbar = xbar
        self.ontarget_check = True

    def h(self, x, t=0):
        return 0

    def g(self, x, u, t=0):
        dJ = 1
        if self.ontarget_check and np.linalg.norm(x - self.xbar) < self.EPS:
            dJ = 0
        return dJ


class QuadraticCostFunctionWithDomainCheck(CostFunction):
    """Quadratic cost, INF where the state is not allowed (costfunction.py:339-415).  Like the reference class it derives
    from CostFunction and borrows QuadraticCostFunction's constructor."""

    def __init__(self, n, m, isavalidstate):
        QuadraticCostFunction.__init__(self, n, m)
        self.isavalidstate = isavalidstate

    @classmethod
    def from_sys(cls, sys):
        inst = cls(sys.n, sys.m, sys.isavalidstate)
        inst.xbar = sys.xbar
        inst.ubar = sys.ubar
        return inst

    def h(self, x, t=0):
        dx = x - self.xbar
        J_f = np.dot(dx.T, np.dot(self.S, dx))
        if not self.isavalidstate(x):
            J_f = self.INF
        if self.ontarget_check and np.linalg.norm(dx) < self.EPS:
            J_f = 0
        return J_f

    def g(self, x, u, t):
        dx = x - self.xbar
        du = u - self.ubar
        dJ = np.dot(dx.T, np.dot(self.Q, dx)) + np.dot(du.T, np.dot(self.R, du))
        if not self.isavalidstate(x):
            dJ = self.INF
        if self.ontarget_check and np.linalg.norm(dx) < self.EPS:
            dJ = 0
        return dJ


class Reachability(CostFunction):
    """g = 0 inside the allowed set (INF outside), h = 0 on the target set (INF elsewhere) (costfunction.py:421-481)."""

    def __init__(self, isavalidestate, xbar=None, isontarget=None):
        super().__init__()
        self.INF = 1E4
        self.EPS = 0.2
        self.isavalidestate = isavalidestate
        if isontarget is None:
            self.isontarget = self.norm_test
            self.xbar = xbar
        else:
            self.isontarget = isontarget

    def norm_test(self, x, t=0):
        return np.linalg.norm(x - self.xbar) < self.EPS

    def h(self, x, t=0):
        return 0 if self.isontarget(x, t) else self.INF

    def g(self, x, u, t=0):
        return 0 if self.isavalidestate(x) else self.INF
