"""Implicit state/action grid — host mirror of pyro.planning.discretizer.GridDynamicSystem.

The reference enumerates every node in Python loops and materialises O(N) / O(N*A) tables
(discretizer.py:167-376); that cannot finish for the 4-D BASELINE grids (1.6e9 nodes, ~50 TB
x_next_table).  This mirror keeps the same public attributes for everything that is O(levels)
or O(A) and turns the O(N) tables into lazily computed, vectorised properties, because the
device kernels work from index arithmetic (node id = C order of x_grid_dim, last axis fastest,
discretizer.py:183-241) and never read them.
"""
import numpy as np
from scipy.interpolate import RegularGridInterpolator


class GridDynamicSystem:
    def __init__(self, sys, x_grid_dim=(101, 101), u_grid_dim=(11,), dt=0.05, lookup=False):
        self.sys = sys
        self.dt = dt
        self.x_grid_dim = np.array(x_grid_dim)
        self.u_grid_dim = np.array(u_grid_dim)
        # the fused kernels need no look-up tables; `lookup=True` builds the reference's dense
        # tables (vectorised) for LUT-mode / small grids only
        self.computelookuptable = lookup
        if sys.n not in (2, 3, 4) or sys.m not in (1, 2):
            raise NotImplementedError  # discretizer.py:243-245, 304-306
        if len(self.x_grid_dim) != sys.n or len(self.u_grid_dim) != sys.m:
            raise ValueError("grid dimensions do not match the system dimensions")
        self.compute()

    # -- discretizer.py:134-163 ----------------------------------------------------------------
    def compute(self):
        self.discretize_state_space()
        if self.computelookuptable:
            # the reference's dense tables (discretizer.py:125-128), only on request: the fused kernels never read them
            self.compute_xnext_table()
            self.compute_action_set_table()

    def discretize_state_space(self):
        """Levels, node / action counts and the O(A) action tables (discretizer.py:134-163, 253-310)."""
        s = self.sys
        self.x_level = [np.linspace(s.x_lb[i], s.x_ub[i], self.x_grid_dim[i]) for i in range(s.n)]
        self.u_level = [np.linspace(s.u_lb[i], s.u_ub[i], self.u_grid_dim[i]) for i in range(s.m)]
        self.nodes_n = int(np.prod(self.x_grid_dim.astype(np.int64)))
        self.actions_n = int(np.prod(self.u_grid_dim.astype(np.int64)))
        self.x_range = s.x_ub - s.x_lb
        self.x_step_size = self.x_range / (self.x_grid_dim - 1)
        self.u_range = s.u_ub - s.u_lb
        self.u_step_size = self.u_range / (self.u_grid_dim - 1)
        # actions are few: enumerate eagerly, C order (discretizer.py:253-310)
        mesh = np.meshgrid(*self.u_level, indexing="ij")
        self.input_from_action_id = np.stack([g.reshape(-1) for g in mesh], axis=1)
        idx = np.meshgrid(*[np.arange(d) for d in self.u_grid_dim], indexing="ij")
        self.index_from_action_id = np.stack([g.reshape(-1) for g in idx], axis=1)
        self.action_id_from_index = np.arange(self.actions_n).reshape(self.u_grid_dim)
        self._state_from_node_id = None
        self._index_from_node_id = None

    # kept for API compatibility: everything they build is built by discretize_state_space / lazily
    def discretize_input_space(self):
        pass

    def generate_nodes(self):
        pass

    def generate_actions(self):
        pass

    # -- dense look-up tables, exactly as the reference builds them (discretizer.py:314-402): O(N*A) host loops,
    #    meant for small grids and for LUT mode with systems the fused kernels do not know ----------------------
    def compute_xnext_table(self):
        if self._device_xnext_table():
            return
        X, U = self.state_from_node_id, self.input_from_action_id
        self.x_next_table = np.zeros((self.nodes_n, self.actions_n, self.sys.n), dtype=float)
        self.x_next_isok = np.zeros((self.nodes_n, self.actions_n), dtype=bool)
        for node_id in range(self.nodes_n):
            x = X[node_id, :]
            for action_id in range(self.actions_n):
                x_next = self.sys.f(x, U[action_id, :]) * self.dt + x
                self.x_next_table[node_id, action_id, :] = x_next
                self.x_next_isok[node_id, action_id] = self.sys.isavalidstate(x_next)

    def _device_xnext_table(self):
        """x_next_table / x_next_isok by pdp_build_tables (one thread per (node, action) pair, the sweep kernels' own
        arithmetic) when the system is one of the four fused ones and a CUDA device is present; False otherwise — the
        tables are a host-side product of the reference API, so the reference's Python loop below remains their
        definition (this is not a sweep path: sweeps never fall back)."""
        try:
            from . import _lib, costfunction, problem
            from .engine import Engine, device_count
            if device_count() == 0:
                return False
            P = problem.extract(self, costfunction.QuadraticCostFunction.from_sys(self.sys))
            if P.system_id == _lib.PDP_SYS_LUT or self.nodes_n * self.actions_n * self.sys.n * 8 > (8 << 30):
                return False
            eng = Engine(P)
        except (RuntimeError, NotImplementedError, ValueError, OSError):
            return False
        try:
            self.x_next_table, self.x_next_isok, _ = eng.build_tables(G=False)
        except NotImplementedError:
            return False
        finally:
            eng.close()
        return True

    def compute_action_set_table(self):
        X, U = self.state_from_node_id, self.input_from_action_id
        self.action_isok = np.zeros((self.nodes_n, self.actions_n), dtype=bool)
        for node_id in range(self.nodes_n):
            for action_id in range(self.actions_n):
                self.action_isok[node_id, action_id] = self.sys.isavalidinput(X[node_id, :], U[action_id, :])

    def compute_nearest_snext_table(self):
        X, U = self.state_from_node_id, self.input_from_action_id
        self.s_next_table = np.zeros((self.nodes_n, self.actions_n), dtype=int)
        for node_id in range(self.nodes_n):
            x = X[node_id, :]
            for action_id in range(self.actions_n):
                x_next = self.sys.f(x, U[action_id, :]) * self.dt + x
                self.s_next_table[node_id, action_id] = self.get_nearest_node_id_from_state(x_next)

    # -- on-disk format of the tables (discretizer.py:412-442): one .npz with the reference's three keys ---------
    def save_lookup_tables(self, name='grid'):
        np.savez(name, x_next_table=self.x_next_table, x_next_isok=self.x_next_isok, action_isok=self.action_isok)

    def load_lookup_tables(self, name='grid'):
        try:
            data = np.load(name + '.npz')
        except Exception:
            print('\n File not found ')
        else:
            self.x_next_table = data['x_next_table']
            self.x_next_isok = data['x_next_isok']
            self.action_isok = data['action_isok']

    # -- O(N) tables, lazily and vectorised (discretizer.py:167-249) -----------------------------
    @property
    def state_from_node_id(self):
        if self._state_from_node_id is None:
            mesh = np.meshgrid(*self.x_level, indexing="ij")
            self._state_from_node_id = np.stack([g.reshape(-1) for g in mesh], axis=1)
        return self._state_from_node_id

    @property
    def index_from_node_id(self):
        if self._index_from_node_id is None:
            idx = np.meshgrid(*[np.arange(d) for d in self.x_grid_dim], indexing="ij")
            self._index_from_node_id = np.stack([g.reshape(-1) for g in idx], axis=1)
        return self._index_from_node_id

    @property
    def node_id_from_index(self):
        return np.arange(self.nodes_n).reshape(self.x_grid_dim)

    # -- conversions (discretizer.py:453-537) ----------------------------------------------------
    def get_index_from_state(self, x):
        return (np.asarray(x, dtype=float) - self.sys.x_lb) / self.x_range * (self.x_grid_dim - 1)

    def get_nearest_index_from_state(self, x):
        return np.clip(np.rint(self.get_index_from_state(x)).astype(int), 0, self.x_grid_dim - 1)

    def get_nearest_node_id_from_state(self, x):
        idx = self.get_nearest_index_from_state(x)
        return int(np.ravel_multi_index(tuple(idx), tuple(self.x_grid_dim)))   # C order = node_id_from_index[idx]

    def get_nearest_index_from_input(self, u):
        return np.clip(np.rint(self.get_index_from_input(u)).astype(int), 0, self.u_grid_dim - 1)

    def get_index_from_input(self, u):
        return (np.asarray(u) - self.sys.u_lb) / self.u_range * (self.u_grid_dim - 1)

    def get_nearest_action_id_from_input(self, u):
        idx = np.clip(np.rint(self.get_index_from_input(u)).astype(int), 0, self.u_grid_dim - 1)
        return self.action_id_from_index[tuple(idx)]

    def get_grid_from_array(self, J):
        return J.reshape(self.x_grid_dim)

    # -- discretizer.py:570-587 -------------------------------------------------------------------
    def compute_interpolation_function(self, J, method="linear", bounds_error=True, fill_value=None):
        if self.nodes_n != J.size:
            raise ValueError("Grid size does not match data")
        levels = tuple(self.x_level[i] for i in range(self.sys.n))
        return RegularGridInterpolator(levels, self.get_grid_from_array(J), method, bounds_error, fill_value)

    def compute_bivariatespline_2D_interpolation_function(self, J, kx=1, ky=1):
        """discretizer.py:592-613 (host-side helper of the spline DP variant, which is not accelerated)."""
        if self.sys.n != 2:
            raise NotImplementedError
        if self.nodes_n != J.size:
            raise ValueError("Grid size does not match data")
        from scipy.interpolate import RectBivariateSpline
        return RectBivariateSpline(self.x_level[0], self.x_level[1], self.get_grid_from_array(J),
                                   bbox=[None, None, None, None], kx=kx, ky=ky)

    def get_2D_slice_of_grid(self, Z, axis_1=0, axis_2=1):
        """2-D slice through the nominal state of an n-D grid array (discretizer.py:637-664)."""
        if self.sys.n == 2:
            return Z
        if self.sys.n < 2:
            raise NotImplementedError
        idx = [int(i) for i in self.get_nearest_index_from_state(self.sys.xbar)]
        idx[axis_1] = slice(None)
        idx[axis_2] = slice(None)
        Z_2d = np.asarray(Z[tuple(idx)], dtype=float)
        return Z_2d if axis_1 < axis_2 else Z_2d.T

    # -- discretizer.py:616-633, vectorised ---------------------------------------------------------
    def get_input_from_policy(self, pi, k):
        if self.nodes_n != pi.size:
            raise ValueError("Grid size does not match optimal action table size")
        return self.input_from_action_id[pi, k]
