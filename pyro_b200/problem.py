"""Problem extraction: pyro objects -> the POD descriptor of include/pyrodp.h.

Reads (lazily, at the first sweep — users mutate ``sys`` bounds before building the grid and
``cf`` / ``dp.alpha`` after building ``dp``; SURVEY.md section 5 "Config / flags"):

  grid_sys.sys (class -> system id, physical constants, x_lb/x_ub/u_lb/u_ub)
  grid_sys.x_grid_dim / u_grid_dim / x_level / u_level / dt      discretizer.py:88-163
  cf class + Q, R, S, xbar, ubar, INF, EPS, ontarget_check         costfunction.py:100-204, 287-334
  dp.alpha                                                          dynamicprogramming.py:130

Works on real pyro objects and on the mirrors in this package alike (duck typing on class
names and attributes).  Transcendentals and matrix inverses are tabulated here per grid level
with the same NumPy calls the reference makes (np.sin / np.cos / np.linalg.inv / np.dot), so
the device never re-derives them and their bits match the reference on the same host.
"""
import ctypes as C

import numpy as np

from . import _lib

_SYS_IDS = {
    "SinglePendulum": _lib.PDP_SYS_PENDULUM,
    "DoublePendulum": _lib.PDP_SYS_TWOLINK,
    "TwoLinkManipulator": _lib.PDP_SYS_TWOLINK,
    "CartPole": _lib.PDP_SYS_CARTPOLE,
}
_COST_IDS = {
    "QuadraticCostFunction": _lib.PDP_COST_QUADRATIC,
    "TimeCostFunction": _lib.PDP_COST_TIME,
    # with the system's own box isavalidstate (checked in classify) a grid node never fails the domain check, so the
    # class IS the quadratic cost on the grid (costfunction.py:339-415)
    "QuadraticCostFunctionWithDomainCheck": _lib.PDP_COST_QUADRATIC,
    "Reachability": _lib.PDP_COST_REACH,
}
_BOX_CHECK_OWNERS = ("ContinuousDynamicSystem", "MechanicalSystem")
# Classes whose methods the fused kernels restate.  A system (cost function) is routed to a fused kernel only if EVERY
# method the sweep depends on is still the one these classes define: a subclass that overrides any of them — the
# reference's own InvertedPendulum flips the sign of g (pendulum.py:283), its Acrobot replaces B (pendulum.py:699) —
# runs in LUT mode on tables built by its own methods instead of silently inheriting its parent's kernel.
_SYS_METHODS = ("f", "ddq", "H", "C", "B", "g", "d", "x2q", "q2x")
_SYS_OWNERS = {
    "SinglePendulum": {"SinglePendulum"},
    "DoublePendulum": {"DoublePendulum", "_TwoLinkForm"},
    "TwoLinkManipulator": {"TwoLinkManipulator", "_TwoLinkForm", "Manipulator"},
    "CartPole": {"CartPole"},
}
_SYS_BASE_OWNERS = {"MechanicalSystem", "ContinuousDynamicSystem"}
_COST_METHODS = ("g", "h", "norm_test")
_COST_BASE_OWNERS = {"CostFunction"}


def _owner(obj, name):
    """Name of the class that defines obj.<name> (None if the attribute is missing or not a plain method)."""
    fn = getattr(type(obj), name, None)
    if fn is None:
        return None
    qual = getattr(fn, "__qualname__", "")
    return qual.split(".")[0] if "." in qual else None


def _class_id(obj, table, methods, owners_of, base_owners):
    """(id, recognised class name) if obj is an instance of a recognised class AND none of `methods` is overridden by a
    class the kernels do not know; (None, None) otherwise."""
    for klass in type(obj).__mro__:
        name = klass.__name__
        if name in table:
            allowed = set(owners_of(name)) | set(base_owners)
            for m in methods:
                own = _owner(obj, m)
                if own is not None and own not in allowed:
                    return None, None
            if any(m in vars(obj) for m in methods):      # a method patched onto the instance
                return None, None
            return table[name], name
    return None, None


def _uses_box_checks(sys):
    """True when isavalidstate / isavalidinput are the base-class box tests (system.py:198-215)."""
    for name in ("isavalidstate", "isavalidinput"):
        if name in vars(sys) or _owner(sys, name) not in _BOX_CHECK_OWNERS:
            return False
    return True


def classify(grid_sys, cf, interpol_method="linear"):
    """Return (system_id, cost_id); system_id == PDP_SYS_LUT means 'needs reference tables'."""
    sys = grid_sys.sys
    sys_id, _ = _class_id(sys, _SYS_IDS, _SYS_METHODS, lambda n: _SYS_OWNERS[n], _SYS_BASE_OWNERS)
    cost_id, _ = _class_id(cf, _COST_IDS, _COST_METHODS, lambda n: {n}, _COST_BASE_OWNERS)
    if interpol_method != "linear":
        raise NotImplementedError("only interpol_method='linear' is accelerated (dynamicprogramming.py:131)")
    if sys_id is None or cost_id is None or not _uses_box_checks(sys) or not _cost_callbacks_are_the_box(sys, cf):
        return _lib.PDP_SYS_LUT, 0
    return sys_id, cost_id


def _cost_callbacks_are_the_box(sys, cf):
    """The cost classes that take callbacks (costfunction.py:339-481) are fused only when the callbacks are the grid system's
    own box test / the class's default norm test; an obstacle function or a custom target set runs in LUT mode."""
    name = type(cf).__name__
    if name not in ("QuadraticCostFunctionWithDomainCheck", "Reachability") and not any(
            k.__name__ in ("QuadraticCostFunctionWithDomainCheck", "Reachability") for k in type(cf).__mro__):
        return True
    valid = getattr(cf, "isavalidstate", None) or getattr(cf, "isavalidestate", None)
    if getattr(valid, "__self__", None) is not sys or getattr(valid, "__func__", None) is not getattr(type(sys), "isavalidstate", None):
        return False
    if hasattr(cf, "isontarget"):
        target = cf.isontarget
        if getattr(target, "__self__", None) is not cf or getattr(target, "__func__", None) is not getattr(type(cf), "norm_test", None):
            return False
        if getattr(cf, "xbar", None) is None:
            return False
    return True


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Problem:
    """Owns the NumPy arrays behind a ``pdp_problem`` struct (kept alive until pdp_create copies)."""

    def __init__(self):
        self.c = _lib.pdp_problem()
        self.keep = []
        self.tables = {}

    def fingerprint(self):
        """Hash of every value the device state depends on (pointer fields excluded)."""
        c = self.c
        scal = [c.n, c.m, c.system_id, c.cost_id, c.ontarget_check, c.slab_begin, c.slab_end, c.alloc_planes,
                list(c.dims), list(c.udims), list(c.x_lb), list(c.x_ub), c.dt, c.alpha, c.INF, c.EPS,
                list(c.Q), list(c.S), list(c.xbar), list(c.sys_par)]
        parts = [repr(scal).encode()] + [self.tables[k].tobytes() for k in sorted(self.tables)]
        return hash(b"".join(parts))

    def hold(self, name, arr):
        arr = _f64(arr)
        self.keep.append(arr)
        self.tables[name] = arr
        return arr


def system_tables(sys, sys_id, x_level):
    """Per-level tables of the state-only transcendental / LAPACK terms (see include/pyrodp.h)."""
    tabs, par = [], np.zeros(8)
    if sys_id == _lib.PDP_SYS_PENDULUM:
        q = x_level[0]
        # g(q) = m1*gravity*lc1*sin(q) (pendulum.py:136), evaluated for all levels at once
        tabs.append(sys.m1 * sys.gravity * sys.lc1 * np.sin(q))
        par[0] = np.linalg.inv(np.asarray(sys.H(np.array([q[0]])), dtype=float))[0, 0]
        par[1] = sys.d1
    elif sys_id == _lib.PDP_SYS_TWOLINK:
        q0, q1 = x_level[0], x_level[1]
        Hinv = np.stack([np.linalg.inv(sys.H(np.array([0.0, a]))) for a in q1])
        tabs.append(Hinv.reshape(-1))
        tabs.append(sys.m2 * sys.l1 * sys.lc2 * np.sin(q1))  # h (pendulum.py:434)
        s1 = np.sin(q0)[:, None]
        s12 = np.sin(q0[:, None] + q1[None, :])
        g1 = (sys.m1 * sys.lc1 + sys.m2 * sys.l1) * sys.gravity
        g2 = sys.m2 * sys.lc2 * sys.gravity
        G = np.stack([-g1 * s1 - g2 * s12, -g2 * s12], axis=-1)  # pendulum.py:465-473
        tabs.append(G.reshape(-1))
        par[0], par[1] = sys.d1, sys.d2
    elif sys_id == _lib.PDP_SYS_CARTPOLE:
        th = x_level[1]
        Hinv = np.stack([np.linalg.inv(sys.H(np.array([0.0, a]))) for a in th])
        tabs.append(Hinv.reshape(-1))
        tabs.append(-sys.m2 * sys.lcg * np.sin(th))           # cartpole.py:399
        tabs.append(sys.m2 * sys.gravity * sys.lcg * np.sin(th))  # cartpole.py:426
    return [_f64(t) for t in tabs], par


def plant_parameters(sys, sys_id):
    """Raw physical parameters of a fused plant in the layout pdp_rollout reads (include/pyrodp.h): what the device needs
    to evaluate f(x, u) at states that are not grid levels (closed-loop rollouts)."""
    ph = np.zeros(16)
    if sys_id == _lib.PDP_SYS_PENDULUM:
        ph[:5] = [sys.m1, sys.lc1, sys.I1, sys.gravity, sys.d1]
    elif sys_id == _lib.PDP_SYS_TWOLINK:
        ph[:10] = [sys.m1, sys.l1, sys.lc1, sys.I1, sys.m2, sys.lc2, sys.I2, sys.gravity, sys.d1, sys.d2]
    elif sys_id == _lib.PDP_SYS_CARTPOLE:
        ph[:4] = [sys.m1, sys.m2, sys.lcg, sys.gravity]
    else:
        raise NotImplementedError("closed-loop rollouts on the device need a fused plant (SinglePendulum, DoublePendulum, "
                                  "TwoLinkManipulator, CartPole); simulate other systems with the reference's simulator")
    return ph


def extract(grid_sys, cf, alpha=1.0, interpol_method="linear", slab=None, alloc_planes=0, force_lut=False,
            lut_actions=None):
    """Build the descriptor for ``pdp_create``.  ``slab`` = (begin, end) axis-0 planes of this rank.
    ``lut_actions``: LUT-mode descriptor with that many table columns per node instead of the grid's
    action count (policy evaluation sweeps have exactly one, dynamicprogramming.py:683-752)."""
    sys = grid_sys.sys
    n, m = int(sys.n), int(sys.m)
    if n not in (2, 3, 4) or m not in (1, 2):
        raise NotImplementedError("grid DP supports n in {2,3,4}, m in {1,2} (discretizer.py:243-245,304-306)")
    sys_id, cost_id = classify(grid_sys, cf, interpol_method)
    if force_lut or lut_actions is not None:
        sys_id, cost_id = _lib.PDP_SYS_LUT, 0

    P = Problem()
    c = P.c
    c.abi_version = _lib.PDP_ABI_VERSION
    c.n, c.m, c.system_id, c.cost_id = n, m, sys_id, cost_id
    dims = [int(d) for d in grid_sys.x_grid_dim]
    udims = [int(d) for d in grid_sys.u_grid_dim]
    if len(dims) != n or len(udims) != m:
        raise ValueError("grid dimensions do not match the system dimensions")
    u_src = grid_sys.u_level
    if lut_actions is not None:
        udims = [int(lut_actions)] + [1] * (m - 1)
        u_src = [np.zeros(d) for d in udims]   # placeholders: the tables carry the inputs' effect
    begin, end = (0, dims[0]) if slab is None else slab
    c.slab_begin, c.slab_end, c.alloc_planes = int(begin), int(end), int(alloc_planes)
    x_level = [P.hold(f"x_level{i}", grid_sys.x_level[i]) for i in range(n)]
    u_level = [P.hold(f"u_level{i}", u_src[i]) for i in range(m)]
    for i in range(n):
        if x_level[i].size != dims[i]:
            raise ValueError("x_level size does not match x_grid_dim")
        c.dims[i] = dims[i]
        c.x_level[i] = _ptr(x_level[i])
        c.x_lb[i], c.x_ub[i] = float(sys.x_lb[i]), float(sys.x_ub[i])
    for i in range(m):
        c.udims[i] = udims[i]
        c.u_level[i] = _ptr(u_level[i])
    c.dt, c.alpha = float(grid_sys.dt), float(alpha)
    c.INF = float(cf.INF)
    c.EPS = float(getattr(cf, "EPS", 0.0))
    c.ontarget_check = int(bool(getattr(cf, "ontarget_check", False)))
    P.A = int(np.prod(udims))
    P.N = int(np.prod(np.array(dims, dtype=np.int64)))
    P.dims, P.udims, P.n, P.m = dims, udims, n, m
    P.system_id, P.cost_id = sys_id, cost_id
    if sys_id == _lib.PDP_SYS_LUT:
        return P

    # ---- cost parameters ----
    xbar = _f64(cf.xbar)
    if xbar.size != n:
        raise ValueError("cf.xbar size does not match the state dimension")
    for i in range(n):
        c.xbar[i] = xbar[i]
    Q = np.zeros((n, n))
    S = np.zeros((n, n))
    if cost_id == _lib.PDP_COST_QUADRATIC:
        Q, S = _f64(cf.Q), _f64(cf.S)
        if Q.shape != (n, n) or S.shape != (n, n):
            raise ValueError("cf.Q / cf.S shape does not match the state dimension")
    for i in range(n):
        for j in range(n):
            c.Q[i * n + j] = Q[i, j]
            c.S[i * n + j] = S[i, j]

    # ---- per-action tables (A is small: plain loops with the reference's own np.dot calls) ----
    mesh = np.meshgrid(*u_level, indexing="ij")
    U = np.stack([g.reshape(-1) for g in mesh], axis=1)  # input_from_action_id, C order
    q_any = np.array([x_level[i][0] for i in range(n // 2)])
    x_any = np.array([x_level[i][0] for i in range(n)])
    B = np.asarray(sys.B(q_any), dtype=float)
    bu = np.stack([np.dot(B, U[a]) for a in range(P.A)])                      # mechanical.py:231
    if cost_id == _lib.PDP_COST_QUADRATIC:
        ubar = _f64(cf.ubar)
        R = _f64(cf.R)
        gu = np.array([np.dot((U[a] - ubar).T, np.dot(R, U[a] - ubar)) for a in range(P.A)])  # costfunction.py:191
    else:
        gu = np.zeros(P.A)
    ok = np.array([1 if sys.isavalidinput(x_any, U[a]) else 0 for a in range(P.A)], dtype=np.uint8)
    bu, gu = P.hold("bu", bu.reshape(-1)), P.hold("gu", gu)
    ok = np.ascontiguousarray(ok)
    P.keep.append(ok)
    P.tables["act_ok"] = ok
    c.bu, c.gu = _ptr(bu), _ptr(gu)
    c.act_ok = ok.ctypes.data_as(C.POINTER(C.c_uint8))

    # ---- state-only system tables ----
    tabs, par = system_tables(sys, sys_id, x_level)
    for i, t in enumerate(tabs):
        P.hold(f"sys_tab{i}", t)
        c.sys_tab[i] = _ptr(P.keep[-1])
        c.sys_tab_len[i] = t.size
    for i in range(8):
        c.sys_par[i] = par[i]
    P.tables["sys_par"] = par
    return P
