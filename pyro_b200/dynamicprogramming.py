"""Drop-in value-iteration planner: the reference's DynamicProgramming API over the B200 engine.

Mirrors pyro/planning/dynamicprogramming.py:
  DynamicProgramming.__init__(grid_sys, cost_function, final_time=0)          :119
  initialize_backward_step / compute_backward_step / finalize_backward_step   :175 / :195 / :240
  compute_steps(n, animate_cost2go, animate_policy, k)                         :265
  solve_bellman_equation(tol, animate_cost2go, animate_policy, k)              :283
  clean_infeasible_set(tol)                                                    :322
  get_lookup_table_controller()                                                :472
  save_latest(name) / load_J_next(name)                                        :481 / :489
  DynamicProgrammingWithLookUpTable                                            :505
  PolicyEvaluator / PolicyEvaluatorWithLookUpTable                             :619 / :677
  LookUpTableController                                                        :27

State visible after any sweep is the reference's: ``J`` (N,) float64, ``pi`` (N,) int64,
``J_next``, ``t``, ``k``, ``alpha``, ``cf``, ``grid_sys``, ``sys`` and, if ``save_time_history``,
``J_list / pi_list / t_list``.  J / pi live on the device; the attributes are fetched lazily.

The sweep itself runs in ``libpyrodp.so`` (CUDA, sm_100a).  There is no CPU path here.
"""
import time

import numpy as np

from . import _lib, problem as _problem
from .engine import Engine, MultiEngine, device_count

# above this many nodes the per-sweep J/pi history (one full array each per sweep,
# dynamicprogramming.py:255-258) defaults to off — 13 GB/sweep at 201^4
HISTORY_MAX_NODES = 1 << 22
# from this many nodes on a fused system is spread over every visible GPU by one host thread (pdp_multi_*)
MULTI_MIN_NODES = 1 << 25


class LookUpTableController:
    """pi -> u(x) by n-linear interpolation of the per-axis input tables (dynamicprogramming.py:27-107), with the
    StaticController surface the reference's controller inherits (pyro/control/controller.py:22-162): dimensions k/m/p,
    name, ref_label/ref_units, r_lb/r_ub, rbar, c / cbar / t2r, forward_kinematic_lines_plus and ``ctl + sys``.
    ``u_tables`` (one (N,) array per input axis, from pdp_get_input_from_policy) replaces the reference's O(N) Python
    loop over get_input_from_policy (discretizer.py:616-633)."""

    def __init__(self, grid_sys, pi, u_tables=None):
        if grid_sys.nodes_n != pi.size:
            raise ValueError("Grid size does not match optimal action table size")
        self.k, self.m, self.p = 1, grid_sys.sys.m, grid_sys.sys.n
        self.grid_sys, self.pi = grid_sys, pi
        self.name = "Tabular Controller"
        self.ref_label = ["Ref. %d" % i for i in range(self.k)]
        self.ref_units = [""] * self.k
        self.r_ub = np.zeros(self.k) + 10
        self.r_lb = np.zeros(self.k) - 10
        self.rbar = np.zeros(self.k)
        self.interpol_method = ["linear"] * self.m
        self._u_tables = u_tables
        self.compute_interpol_functions()

    def compute_interpol_functions(self):
        self.u_interpol = []
        for k in range(self.m):
            u_k = self._u_tables[k] if self._u_tables is not None else self.grid_sys.get_input_from_policy(self.pi, k)
            self.u_interpol.append(self.grid_sys.compute_interpolation_function(
                u_k, self.interpol_method[k], bounds_error=False, fill_value=0))

    def lookup_table_selection(self, x):
        u = np.zeros(self.m)
        for k in range(self.m):
            u[k] = self.u_interpol[k](x)[0]
        return u

    def c(self, y, r, t=0):
        return self.lookup_table_selection(y)

    # StaticController (pyro/control/controller.py:93-162): constant reference, feedback at it, closed loop by "+"
    def t2r(self, t):
        return self.rbar

    def cbar(self, y, t=0):
        return self.c(y, self.t2r(t), t)

    def forward_kinematic_lines_plus(self, x, u, t):
        return None, None, None

    def __add__(self, sys):
        """closed_loop_system = controller + dynamic_system (controller.py:155-162): pyro's own ClosedLoopSystem."""
        try:
            from pyro.control.controller import ClosedLoopSystem
        except Exception as exc:
            raise ImportError("ctl + sys builds pyro.control.controller.ClosedLoopSystem; the pyro package is not importable "
                              "here (closed-loop simulation is outside the accelerated path)") from exc
        return ClosedLoopSystem(sys, self)

    def plot_control_law(self, *a, **k):
        raise NotImplementedError("plotting is outside the accelerated path; wrap the tables in pyro's own controller "
                                  "(make_reference_controller) to use its plot helpers")


def make_reference_controller(grid_sys, pi, u_tables):
    """The reference's own LookUpTableController (a pyro StaticController with every plot / closed-loop helper) fed with
    device-computed input tables, when the pyro package is importable; None otherwise."""
    try:
        from pyro.planning.dynamicprogramming import LookUpTableController as RefController
    except Exception:
        return None

    class DeviceLookUpTableController(RefController):
        def compute_interpol_functions(self):      # dynamicprogramming.py:72-83 without the per-node Python loop
            self.u_interpol = [self.grid_sys.compute_interpolation_function(u_tables[k], self.interpol_method[k],
                                                                            bounds_error=False, fill_value=0)
                               for k in range(self.m)]

    return DeviceLookUpTableController(grid_sys, pi)


class DynamicProgramming:
    """Dynamic programming on a grid sys — Bellman sweeps on the GPU."""

    def __init__(self, grid_sys, cost_function, final_time=0, engine_factory=None, shard=False, time_varying=False):
        self.grid_sys = grid_sys
        self.sys = grid_sys.sys
        self.cf = cost_function
        self.tf = final_time
        self.alpha = 1.0
        self.interpol_method = "linear"
        self.save_time_history = grid_sys.nodes_n <= HISTORY_MAX_NODES
        self.max_sweeps = None  # optional guard for solve_bellman_equation (hard part 7)
        self.shard = shard      # True / a process group: slabs of the grid over the ranks of torch.distributed (one GPU each)
        self.time_varying = time_varying  # table fallback only: rebuild x_next / G at every sweep's t, as the base class re-evaluates
                                   # f(x,u,t) and g(x,u,t) (dynamicprogramming.py:214,223); the table class freezes them at tf
        self.verbose = True
        self.t = self.tf
        self.k = 0
        self.start_time = time.time()
        self._engine_factory = engine_factory
        self._engine = None
        self._key = None
        self._J = self._pi = self._J_next = None
        self._last_stats = None
        self.evaluate_terminal_cost()
        if self.save_time_history:
            self.t_list, self.J_list, self.pi_list = [self.tf], [self.J], [self.pi]

    # ---- engine management -----------------------------------------------------------------------
    def _extract(self):
        return _problem.extract(self.grid_sys, self.cf, self.alpha, self.interpol_method)

    def _make_engine(self, P):
        if self._engine_factory is not None:
            return self._make_lut_engine(P) if P.system_id == _lib.PDP_SYS_LUT else self._engine_factory(self, P)
        if self.shard:
            # one process per GPU (torchrun): opt-in, because every rank must then build and drive the same planner —
            # get_J / get_pi / parameter changes become collectives
            from . import distributed
            if not distributed.is_sharded():
                raise RuntimeError("dp.shard is set but torch.distributed is not initialised with more than one rank")
            if P.system_id == _lib.PDP_SYS_LUT:
                raise NotImplementedError("sharded sweeps need a fused system (SinglePendulum, DoublePendulum, TwoLinkManipulator, "
                                          "CartPole with box bounds): an arbitrary x_next_table has no a-priori halo")
            group = None if self.shard is True else self.shard
            return distributed.ShardedEngine(self.grid_sys, self.cf, self.alpha, self.interpol_method, group=group)
        if P.system_id == _lib.PDP_SYS_LUT:
            return self._make_lut_engine(P)
        n_parts = self._multi_parts(P)
        if n_parts > 1:
            try:
                return MultiEngine(P, n_parts=n_parts)
            except ValueError:          # the halo is wider than a slab of this grid: one handle
                pass
        return Engine(P)

    @staticmethod
    def _multi_parts(P):
        """How many slab handles this process drives (pdp_multi_*): every visible GPU for a grid large enough to pay for
        the halo copies, one otherwise.  PYRODP_MULTI = 0 / 1 (off), n (that many parts, round-robin over the GPUs)."""
        import os
        env = os.environ.get("PYRODP_MULTI")
        if env is not None:
            return max(int(env), 1)
        ndev = device_count()
        return ndev if (ndev > 1 and P.N >= MULTI_MIN_NODES) else 1

    # Which INF semantics the table fallback reproduces.  The reference's base class gives a disallowed input exactly INF
    # (dynamicprogramming.py:230-233); its table class computes G + alpha*J(x_next) with G = INF, i.e. INF + alpha*J when the
    # arrival state lies inside the grid (:545-549, :567).  Both are reproduced (the PolicyEvaluator pair does the same).
    _invalid_input_is_exact_inf = True

    def _make_lut_engine(self, P):
        eng = self._engine_factory(self, P) if self._engine_factory is not None else Engine(P)
        x_next, G = build_lookup_tables(self.grid_sys, self.cf, self.t if self.time_varying else self.tf,
                                        exact_inf=self._invalid_input_is_exact_inf, use_grid_tables=not self.time_varying)
        eng.set_lut(x_next, G)
        return eng

    def _ensure_engine(self):
        """(Re)build the device state if any parameter the sweep reads has changed."""
        P = self._extract()
        key = P.fingerprint()
        if self._engine is None or key != self._key:
            carry = None
            if self._engine is not None:
                carry = self._engine.get_J()
                self._engine.close()
            self._engine = self._make_engine(P)
            self._key = key
            if carry is not None:
                self._engine.set_J(carry)
        return self._engine

    # ---- lazily fetched arrays ----------------------------------------------------------------------
    @property
    def J(self):
        if self._J is None:
            self._J = self._engine.get_J()
        return self._J

    @J.setter
    def J(self, value):
        self._J = np.array(value, dtype=float)
        self._ensure_engine().set_J(self._J)

    @property
    def pi(self):
        if self._pi is None:
            self._pi = self._engine.get_pi()
        return self._pi

    @pi.setter
    def pi(self, value):
        self._pi = np.asarray(value)

    @property
    def J_next(self):
        if self._J_next is None:
            self._J_next = self._engine.get_J_next()
        return self._J_next

    @J_next.setter
    def J_next(self, value):
        self._J_next = value

    def _invalidate(self):
        self._J = self._pi = self._J_next = None

    # ---- reference hooks ---------------------------------------------------------------------------
    def evaluate_terminal_cost(self):
        """J = cf.h(x, tf) on every node, pi = 0 (dynamicprogramming.py:159-171), on the device."""
        eng = self._ensure_engine()
        if getattr(eng, "lut_mode", False) or eng.problem.system_id == _lib.PDP_SYS_LUT:
            xs = self.grid_sys.state_from_node_id
            eng.set_J(np.array([self.cf.h(xs[s, :], self.tf) for s in range(self.grid_sys.nodes_n)], dtype=float))
        else:
            eng.eval_terminal_cost()
        self._invalidate()
        self._pi = np.zeros(self.grid_sys.nodes_n, dtype=int)

    def initialize_backward_step(self):
        self.k = self.k + 1
        self.t = self.t - self.grid_sys.dt
        eng = self._ensure_engine()
        if self.time_varying and eng.problem.system_id == _lib.PDP_SYS_LUT and hasattr(eng, "set_lut"):
            x_next, G = build_lookup_tables(self.grid_sys, self.cf, self.t, exact_inf=self._invalid_input_is_exact_inf,
                                            use_grid_tables=False)
            eng.set_lut(x_next, G)

    def compute_backward_step(self):
        self._last_stats = self._engine.sweep(1)[0]
        self._invalidate()

    def finalize_backward_step(self):
        return self._report(self._last_stats, self.k, self.t)

    def _report(self, stats, k, t):
        elapsed_time = time.time() - self.start_time
        j_max, delta_max, delta_min = (float(v) for v in stats)
        if self.verbose:
            print('%d t:%.2f Elasped:%.2f max: %.2f dmax:%.2f dmin:%.2f' % (k, t, elapsed_time, j_max, delta_max, delta_min))
        if self.save_time_history:
            self.J_list.append(self.J)
            self.t_list.append(t)
            self.pi_list.append(self.pi)
        return abs(np.array([delta_max, delta_min])).max()

    # ---- drivers -----------------------------------------------------------------------------------
    def compute_steps(self, n=50, animate_cost2go=False, animate_policy=False, k=0):
        if animate_cost2go or animate_policy:
            raise NotImplementedError("plot/animation helpers are outside the accelerated path")
        if self.verbose:
            print('\nComputing %d backward DP iterations:' % n)
            print('-----------------------------------------')
        if self.save_time_history:
            for _ in range(n):
                self.initialize_backward_step()
                self.compute_backward_step()
                self.finalize_backward_step()
            return
        # no per-sweep host copies wanted: run all n sweeps back to back on the device
        eng = self._ensure_engine()
        stats = eng.sweep(n)
        self._invalidate()
        for i in range(n):
            self.k += 1
            self.t = self.t - self.grid_sys.dt
            self._report(stats[i], self.k, self.t)
        self._last_stats = stats[-1] if n else self._last_stats

    def solve_bellman_equation(self, tol=0.1, animate_cost2go=False, animate_policy=False, k=0):
        if animate_cost2go or animate_policy:
            raise NotImplementedError("plot/animation helpers are outside the accelerated path")
        if self.verbose:
            print('\nComputing backward DP iterations until dJ<%2.2f:' % tol)
            print('---------------------------------------------------------')
        delta = self.cf.INF
        sweeps = 0
        while delta > tol:
            if self.max_sweeps is not None and sweeps >= self.max_sweeps:
                if self.verbose:
                    print('\nmax_sweeps reached before dJ<tol')
                return
            self.initialize_backward_step()
            self.compute_backward_step()
            delta = self.finalize_backward_step()
            sweeps += 1
        if self.verbose:
            print('\nBellman equation solved!')

    # ---- data tools ----------------------------------------------------------------------------------
    def clean_infeasible_set(self, tol=1):
        default_action = self.grid_sys.get_nearest_action_id_from_input(self.sys.ubar)
        self._engine.clean_infeasible_set(tol, int(default_action))
        self._J = self._pi = None

    def get_lookup_table_controller(self):
        """dynamicprogramming.py:472-477.  With the pyro package importable this IS pyro's LookUpTableController (so
        ``ctl + sys``, ``plot_control_law`` ... behave as in the reference); otherwise the mirror above."""
        u_tables = [self._engine.get_input_from_policy(k) for k in range(self.sys.m)]
        ctl = make_reference_controller(self.grid_sys, self.pi, u_tables) if hasattr(self.grid_sys, "x_grid_dim") else None
        return ctl if ctl is not None else LookUpTableController(self.grid_sys, self.pi, u_tables)

    def compute_closed_loop_trajectories(self, x0, tf=10, n=10001, stride=1):
        """Closed-loop Euler trajectories of the plant under the current policy, for a BATCH of initial states on the
        device: what ``cl_sys = dp.get_lookup_table_controller() + sys; cl_sys.x0 = x0; cl_sys.compute_trajectory(tf, n,
        'euler')`` computes for one (simulation.py:298-324, controller.py:326-355, dynamicprogramming.py:85-107).
        x0 (B, n_states) -> t (n_keep,), x (B, n_keep, n_states), u (B, n_keep, m); every ``stride``-th point is kept."""
        eng = self._engine
        if eng is None or not hasattr(eng, "rollout"):
            raise NotImplementedError("closed-loop rollouts need the policy on one device handle (not a sharded / multi-part run)")
        if int(n) < 2 or not tf > 0:
            raise ValueError("compute_closed_loop_trajectories: needs n >= 2 points and tf > 0")
        phys = _problem.plant_parameters(self.sys, eng.problem.system_id)
        dt = (tf + 0.0 - 0) / (n - 1)
        x, u = eng.rollout(phys, x0, n, dt, stride)
        return np.linspace(0, tf, n)[::stride], x, u

    def save_latest(self, name='test_data'):
        np.save(name + '_J_inf', self.J_next)
        np.save(name + '_pi_inf', self.pi.astype(int))

    def load_J_next(self, name='test_data'):
        try:
            self.J_next = np.load(name + '_J_inf' + '.npy')
        except Exception:
            print('Failed to load J_next ')

    def plot_cost2go(self, *a, **k):
        raise NotImplementedError("plotting is outside the accelerated path; use the reference's helpers on dp.J")

    plot_policy = plot_cost2go_3D = animate_cost2go = animate_policy = plot_cost2go


class DynamicProgrammingWithLookUpTable(DynamicProgramming):
    """Name kept for drop-in use: every reference example instantiates this class
    (dynamicprogramming.py:505).  Known systems run the fused on-the-fly kernel (no tables are
    ever materialised); anything else runs the LUT-mode kernel on the reference-style tables, with the table class's
    own value INF + alpha*J(x_next) where only the input is disallowed (:545-549, :567)."""
    _invalid_input_is_exact_inf = False


class DynamicProgramming2DRectBivariateSpline(DynamicProgrammingWithLookUpTable):
    """dynamicprogramming.py:578-614: the table sweep with J_next interpolated by scipy's RectBivariateSpline(kx=3, ky=3)
    — the interpolating bicubic spline, arguments outside the grid clamped to its edge — instead of the
    RegularGridInterpolator.  2-D grids only (discretizer.py:600-612 raises NotImplementedError otherwise).  Always table
    mode, as in the reference (Q = G + alpha * J_interpol(x_next_table)); the spline is refitted to J_next on the device
    before every backup (pdp_set_interpolant).  Floating-point parity with the reference, not bit parity."""

    def _extract(self):
        if self.sys.n != 2:
            raise NotImplementedError("the bivariate-spline interpolant exists for 2-D grids only (discretizer.py:600-612)")
        return _problem.extract(self.grid_sys, self.cf, self.alpha, self.interpol_method, force_lut=True)

    def _fused_twin(self):
        """Descriptor of the same problem on a fused kernel, or None: lets the tables and the terminal cost of a plant the
        library knows come from the device (pdp_build_tables, pdp_eval_terminal_cost) instead of O(N*A) Python loops."""
        if self._engine_factory is not None or self.time_varying:
            return None
        P = _problem.extract(self.grid_sys, self.cf, self.alpha, self.interpol_method)
        return None if P.system_id == _lib.PDP_SYS_LUT else P

    def _make_engine(self, P):
        if self._engine_factory is not None:
            eng = self._engine_factory(self, P)
            eng.set_lut(*build_lookup_tables(self.grid_sys, self.cf, self.tf, exact_inf=False))
        else:
            twin = self._fused_twin()
            if twin is None:
                eng = self._make_lut_engine(P)
            else:
                src = Engine(twin)
                try:
                    x_next, _, G = src.build_tables(x_ok=False)      # G = g*dt where input and arrival state are allowed, else INF
                finally:
                    src.close()
                eng = Engine(P)
                eng.set_lut(x_next, G)
        eng.set_interpolant("spline3")
        return eng

    def evaluate_terminal_cost(self):
        twin = self._fused_twin()
        if twin is None:
            return super().evaluate_terminal_cost()
        src = Engine(twin)
        try:
            src.eval_terminal_cost()
            J = src.get_J()
        finally:
            src.close()
        self._ensure_engine().set_J(J)
        self._invalidate()
        self._pi = np.zeros(self.grid_sys.nodes_n, dtype=int)


class PolicyEvaluator(DynamicProgramming):
    """Evaluate the cost-to-go of a given control law (dynamicprogramming.py:619-672): the backup
    J[s] = g(x_s, u_s)*dt + alpha*J_next(x_next_s) with u_s = ctl.c(x_s, ctl.rbar, t), INF where the
    input or the arrival state is not allowed.  One column per node, no min: the tables of
    PolicyEvaluatorWithLookUpTable (:683-729) are built once on the host exactly as the reference builds
    them and the sweeps run on the device in LUT mode (J = G + alpha*RGI(J_next)(x_next_table), :743-752)."""

    def __init__(self, ctl, grid_sys, cost_function, final_time=0, engine_factory=None):
        self.ctl = ctl
        DynamicProgramming.__init__(self, grid_sys, cost_function, final_time, engine_factory)

    def _extract(self):
        return _problem.extract(self.grid_sys, self.cf, self.alpha, self.interpol_method, lut_actions=1)

    def _make_engine(self, P):
        if self._engine_factory is not None:
            return self._engine_factory(self, P)
        eng = Engine(P)
        self.compute_lookuptable()
        eng.set_lut(self.x_next_table, self.G)
        return eng

    # Base class semantics (dynamicprogramming.py:636-672): a node whose input OR arrival state is not allowed gets
    # exactly INF.  The table variant (:700-752) instead computes INF + alpha*J(x_next) when only the input is
    # disallowed and x_next lies inside the grid.  Both are reproduced: here such a node's table entry is moved
    # outside the box, where the interpolation returns its fill value 0.
    _invalid_input_is_exact_inf = True

    def compute_lookuptable(self):
        """x_next_table (N, n) and G (N,) of the control law (dynamicprogramming.py:683-729)."""
        gs, sys, cf = self.grid_sys, self.sys, self.cf
        X = gs.state_from_node_id
        self.x_next_table = np.zeros((gs.nodes_n, sys.n), dtype=float)
        self.G = np.zeros(gs.nodes_n, dtype=float)
        outside = np.asarray(sys.x_ub, dtype=float) + 1.0
        for s in range(gs.nodes_n):
            x = X[s, :]
            u = self.ctl.c(x, self.ctl.rbar, self.t)
            x_next = sys.f(x, u, self.t) * gs.dt + x
            self.x_next_table[s, :] = x_next
            u_ok = sys.isavalidinput(x, u)
            if u_ok and sys.isavalidstate(x_next):
                self.G[s] = cf.g(x, u, self.t) * gs.dt
            else:
                self.G[s] = cf.INF
                if not u_ok and self._invalid_input_is_exact_inf:
                    self.x_next_table[s, :] = outside

    def get_lookup_table_controller(self):
        raise NotImplementedError("a policy evaluation has no policy table; the controller is self.ctl")

    def clean_infeasible_set(self, tol=1):
        raise NotImplementedError("clean_infeasible_set rewrites the policy; a policy evaluation has none")


class PolicyEvaluatorWithLookUpTable(PolicyEvaluator):
    """The table variant (dynamicprogramming.py:677-752), with the reference's own tables and its INF + alpha*J
    value on nodes whose input alone is disallowed."""
    _invalid_input_is_exact_inf = False


def build_lookup_tables(grid_sys, cf, t=0, exact_inf=False, use_grid_tables=True):
    """Reference-style dense tables for LUT mode: x_next (N,A,n), G (N,A) with INF folded in
    (discretizer.py:342-376, dynamicprogramming.py:517-553).  Uses the grid's own tables when it has them (a real pyro
    GridDynamicSystem built with lookup=True; they are evaluated at the default t), else calls sys.f(x, u, t).
    ``exact_inf``: base-class semantics — a disallowed input costs exactly INF (dynamicprogramming.py:230-233): its table
    entry is moved outside the box, where the interpolation returns its fill value 0."""
    sys = grid_sys.sys
    N, A, n = grid_sys.nodes_n, grid_sys.actions_n, sys.n
    X, U = grid_sys.state_from_node_id, grid_sys.input_from_action_id
    have = use_grid_tables and all(hasattr(grid_sys, a) for a in ("x_next_table", "x_next_isok", "action_isok"))
    if have:
        x_next, x_ok, a_ok = np.array(grid_sys.x_next_table, dtype=float), grid_sys.x_next_isok, grid_sys.action_isok
    else:
        x_next = np.zeros((N, A, n))
        x_ok = np.zeros((N, A), dtype=bool)
        a_ok = np.zeros((N, A), dtype=bool)
        for s in range(N):
            for a in range(A):
                xn = sys.f(X[s, :], U[a, :], t) * grid_sys.dt + X[s, :]
                x_next[s, a, :] = xn
                x_ok[s, a] = sys.isavalidstate(xn)
                a_ok[s, a] = sys.isavalidinput(X[s, :], U[a, :])
    G = np.full((N, A), float(cf.INF))
    for s in range(N):
        for a in range(A):
            if a_ok[s, a] and x_ok[s, a]:
                G[s, a] = cf.g(X[s, :], U[a, :], t) * grid_sys.dt
    if exact_inf:
        x_next[~np.asarray(a_ok, dtype=bool)] = np.asarray(sys.x_ub, dtype=float) + 1.0
    return x_next, G
