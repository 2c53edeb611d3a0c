"""Host-side parameter holders for the four BASELINE systems.

These mirror the *attribute names and default values* of the reference models so the problem
extractor (``problem.py``) can read either these objects or real pyro objects by duck typing:

  SinglePendulum      pyro/dynamic/pendulum.py:16-150
  DoublePendulum      pyro/dynamic/pendulum.py:340-493
  TwoLinkManipulator  pyro/dynamic/manipulator.py:795-992
  CartPole            pyro/dynamic/cartpole.py:322-437
  base bounds         pyro/dynamic/system.py:90-97, pyro/dynamic/mechanical.py:59-74

They are NOT a re-implementation of pyro's dynamics package: only what the Bellman sweep reads
(dimensions, bounds, physical constants, and the H/C/B/g/d terms of
``H ddq + C dq + d + g = B u`` evaluated per grid level when the device tables are built).
Each term keeps the reference's expression order so that, evaluated with NumPy on the same
host, the table entries have the same bits as the reference's own calls.
"""
import numpy as np


class MechanicalSystem:
    """dof-joint mechanical system, state x = [q, dq] (mechanical.py:36-74)."""

    def __init__(self, dof=1, actuators=None):
        self.dof = dof
        self.n = 2 * dof
        self.m = dof if actuators is None else actuators
        self.p = self.n
        self.name = f"{dof}DoF Mechanical System"
        self.x_ub = np.zeros(self.n) + 2 * np.pi
        self.x_lb = np.zeros(self.n) - 2 * np.pi
        self.u_ub = np.zeros(self.m) + 5
        self.u_lb = np.zeros(self.m) - 5
        self.xbar = np.zeros(self.n)
        self.ubar = np.zeros(self.m)

    # -- terms of the manipulator equation; subclasses override ------------------------------
    def H(self, q):
        return np.diag(np.ones(self.dof))

    def C(self, q, dq):
        return np.zeros((self.dof, self.dof))

    def B(self, q):
        B = np.zeros((self.dof, self.m))
        for i in range(min(self.m, self.dof)):
            B[i, i] = 1
        return B

    def g(self, q):
        return np.zeros(self.dof)

    def d(self, q, dq):
        return np.zeros(self.dof)

    # -- forward dynamics (mechanical.py:222-263) ---------------------------------------------
    def ddq(self, q, dq, u, t=0):
        rhs = np.dot(self.B(q), u) - np.dot(self.C(q, dq), dq) - self.g(q) - self.d(q, dq)
        return np.dot(np.linalg.inv(self.H(q)), rhs)

    def f(self, x, u, t=0):
        q, dq = x[: self.dof], x[self.dof:]
        dx = np.zeros(self.n)
        dx[: self.dof] = dq
        dx[self.dof:] = self.ddq(q, dq, u, t)
        return dx

    # -- box domain checks, strict compares (system.py:198-215) --------------------------------
    def isavalidstate(self, x):
        return not any((x[i] < self.x_lb[i]) or (x[i] > self.x_ub[i]) for i in range(self.n))

    def isavalidinput(self, x, u):
        return not any((u[i] < self.u_lb[i]) or (u[i] > self.u_ub[i]) for i in range(self.m))


class SinglePendulum(MechanicalSystem):
    def __init__(self):
        super().__init__(1)
        self.name = "Single Pendulum"
        self.l1, self.lc1 = 2.0, 1
        self.m1, self.I1, self.gravity, self.d1 = 1, 1, 9.81, 0

    def H(self, q):
        return np.array([[self.m1 * self.lc1 ** 2 + self.I1]], dtype=float)

    def B(self, q):
        return np.diag(np.ones(self.dof))

    def g(self, q):
        return np.array([self.m1 * self.gravity * self.lc1 * np.sin(q[0])])

    def d(self, q, dq):
        return np.array([self.d1 * dq[0]])


class _TwoLinkForm(MechanicalSystem):
    """Shared 2-link H, C, g, d (pendulum.py:400-493 == manipulator.py:897-992)."""

    def H(self, q):
        c2 = np.cos(q[1])
        H = np.zeros((2, 2))
        H[0, 0] = (self.m1 * self.lc1 ** 2 + self.I1
                   + self.m2 * (self.l1 ** 2 + self.lc2 ** 2 + 2 * self.l1 * self.lc2 * c2) + self.I2)
        H[1, 0] = self.m2 * self.lc2 ** 2 + self.m2 * self.l1 * self.lc2 * c2 + self.I2
        H[0, 1] = H[1, 0]
        H[1, 1] = self.m2 * self.lc2 ** 2 + self.I2
        return H

    def coriolis_h(self, q):
        return self.m2 * self.l1 * self.lc2 * np.sin(q[1])

    def C(self, q, dq):
        h = self.coriolis_h(q)
        C = np.zeros((2, 2))
        C[0, 0] = -h * dq[1]
        C[1, 0] = h * dq[0]
        C[0, 1] = -h * (dq[0] + dq[1])
        return C

    def B(self, q):
        return np.diag(np.ones(self.dof))

    def g(self, q):
        s1, s12 = np.sin(q[0]), np.sin(q[0] + q[1])
        g1 = (self.m1 * self.lc1 + self.m2 * self.l1) * self.gravity
        g2 = self.m2 * self.lc2 * self.gravity
        return np.array([-g1 * s1 - g2 * s12, -g2 * s12])

    def d(self, q, dq):
        return np.dot(np.array([[self.d1, 0], [0, self.d2]]), dq)


class DoublePendulum(_TwoLinkForm):
    def __init__(self):
        super().__init__(2)
        self.name = "Double Pendulum"
        self.l1 = self.l2 = self.lc1 = self.lc2 = 1
        self.m1, self.I1, self.m2, self.I2 = 1, 0, 1, 0
        self.gravity = 9.81
        self.d1 = self.d2 = 0


class TwoLinkManipulator(_TwoLinkForm):
    def __init__(self):
        super().__init__(2, 2)
        self.name = "Two Link Manipulator"
        self.l1, self.l2, self.lc1, self.lc2 = 0.5, 0.3, 0.2, 0.1
        self.m1, self.I1, self.m2, self.I2 = 1, 0, 1, 0
        self.gravity = 9.81
        self.d1 = self.d2 = 0.5


class CartPole(MechanicalSystem):
    def __init__(self):
        super().__init__(dof=2, actuators=1)
        self.name = "Cart Pole"
        self.u_lb[0], self.u_ub[0] = -10, +10
        self.l, self.lcg = 3, 0.5
        self.m1, self.m2, self.gravity = 1, 0.1, 9.81

    def H(self, q):
        H = np.zeros((2, 2))
        H[0, 0] = self.m1 + self.m2
        H[1, 0] = self.m2 * self.lcg * np.cos(q[1])
        H[0, 1] = H[1, 0]
        H[1, 1] = self.m2 * self.lcg ** 2
        return H

    def coriolis_k(self, q):
        # C[0,1] = -m2*lcg*sin(theta) * theta_dot (cartpole.py:399); this is the state-only factor
        return -self.m2 * self.lcg * np.sin(q[1])

    def C(self, q, dq):
        C = np.zeros((2, 2))
        C[0, 1] = self.coriolis_k(q) * dq[1]
        return C

    def B(self, q):
        B = np.zeros((2, 1))
        B[0] = 1
        return B

    def g(self, q):
        return np.array([0.0, self.m2 * self.gravity * self.lcg * np.sin(q[1])])


SYSTEMS = {c.__name__: c for c in (SinglePendulum, DoublePendulum, TwoLinkManipulator, CartPole)}
