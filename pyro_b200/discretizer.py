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

    # -- discretizer.py:616-633, vectorised ---------------------------------------------------------
    def get_input_from_policy(self, pi, k):
        if self.nodes_n != pi.size:
            raise ValueError("Grid size does not match optimal action table size")
        return self.input_from_action_id[pi, k]
