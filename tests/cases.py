"""Parity cases shared by the golden generator (reference side) and the tests (our side).

Each case is plain data.  ``build_case`` instantiates it on the pyro_b200 mirrors exactly the
way oracle/gen_golden.py instantiates it on the real reference classes.
"""
import numpy as np

PI = float(np.pi)

CASES = {
    # BASELINE config 1: SinglePendulum 51x51, 11 actions (dynamicprogramming.py:774-783 settings)
    "pend_51x51x11": dict(system="SinglePendulum", x_grid_dim=[51, 51], u_grid_dim=[11], xbar=[-3.14, 0.0], INF=300.0,
                          snapshots=[1, 2, 10, 50, 100]),
    "pend_101x101x21": dict(system="SinglePendulum", x_grid_dim=[101, 101], u_grid_dim=[21], xbar=[-3.14, 0.0], INF=300.0,
                            snapshots=[1, 10, 50], table_stride=7),
    # non-square grid, discount, damping, custom bounds, minimum-time cost with a real target zone
    "pend_time_41x61x7": dict(system="SinglePendulum", x_grid_dim=[41, 61], u_grid_dim=[7], cost="time", xbar=[-3.14, 0.0],
                              INF=50.0, EPS=0.5, alpha=0.9, dt=0.1, sys_params={"d1": 0.3},
                              x_lb=[-4.0, -5.0], x_ub=[1.0, 6.0], u_lb=[-8.0], u_ub=[8.0], snapshots=[1, 5, 20]),
    # the reference's only 4-D example (examples/demos_by_tool/dynamicprogramming/double_pendulum_optimal_swingup.py:19-54)
    "dpend_example": dict(system="DoublePendulum", x_grid_dim=[11, 9, 13, 11], u_grid_dim=[3, 5], dt=0.1,
                          x_lb=[-5.0, -1.5, -4.0, -4.0], x_ub=[0.5, 4.0, 5.5, 7.0], u_lb=[-12.0, -12.0], u_ub=[12.0, 12.0],
                          xbar=[0.0, 0.0, 0.0, 0.0], Q=[1.0, 0.5, 0.1, 0.05], R=[0.05, 0.05], INF=1000.0, EPS=1.0,
                          snapshots=[1, 2, 10, 30], table_stride=5),
    "dpend_default_11": dict(system="DoublePendulum", x_grid_dim=[11, 11, 11, 11], u_grid_dim=[3, 3], INF=1000.0,
                             snapshots=[1, 2, 10], table_stride=5),
    "twolink_9": dict(system="TwoLinkManipulator", x_grid_dim=[9, 9, 9, 9], u_grid_dim=[3, 3], INF=1000.0,
                      snapshots=[1, 2, 10], table_stride=3),
    # small torques so most transitions stay in bounds and the interpolation is exercised
    "twolink_soft": dict(system="TwoLinkManipulator", x_grid_dim=[9, 11, 9, 13], u_grid_dim=[5, 3], INF=500.0, alpha=0.95,
                         u_lb=[-0.3, -0.05], u_ub=[0.3, 0.05], snapshots=[1, 2, 10], table_stride=3),
    "cartpole_swingup": dict(system="CartPole", x_grid_dim=[9, 11, 9, 11], u_grid_dim=[5], xbar=[0.0, PI, 0.0, 0.0],
                             INF=1000.0, snapshots=[1, 2, 10, 30], table_stride=3),
    # the reference's reachability example (examples/demos_by_tool/dynamicprogramming/pendulum_reachability.py:16-27):
    # costfunction.Reachability(sys.isavalidstate, sys.xbar), class defaults INF = 1e4, EPS = 0.2
    "pend_reach_41x41x3": dict(system="SinglePendulum", x_grid_dim=[41, 41], u_grid_dim=[3], cost="reach", xbar=[-3.14, 0.0],
                               snapshots=[1, 5, 30], table_stride=3),
    # QuadraticCostFunctionWithDomainCheck.from_sys (costfunction.py:339-415) on a box-bounded system
    "cartpole_domaincheck": dict(system="CartPole", x_grid_dim=[7, 9, 7, 9], u_grid_dim=[5], cost="domaincheck",
                                 xbar=[0.0, PI, 0.0, 0.0], INF=1000.0, S=[1.0, 2.0, 0.5, 0.1], snapshots=[1, 3, 10], table_stride=3),
}


class LinearFeedback:
    """Static state feedback u = -K (y - ref), the controller interface PolicyEvaluator reads
    (pyro/control/controller.py: ``rbar`` and ``c(y, r, t)``).  Unsaturated on purpose, so that some
    nodes ask for a disallowed input and take the INF branch with an in-bounds arrival state."""

    def __init__(self, K, ref):
        self.K = np.array(K, float)
        self.ref = np.array(ref, float)
        self.rbar = np.zeros(1)

    def c(self, y, r, t=0):
        return -np.dot(self.K, np.asarray(y, float) - self.ref)


# policy evaluation (dynamicprogramming.py:619-752): J of a fixed control law, one table column per node
POLICY_CASES = {
    "pe_pend_pd": dict(system="SinglePendulum", x_grid_dim=[41, 37], u_grid_dim=[3], xbar=[-3.14, 0.0], INF=300.0,
                       ctl=dict(K=[[0.5, 0.3]], ref=[-3.14, 0.0]), snapshots=[1, 5, 40]),
    "pe_cartpole_lin": dict(system="CartPole", x_grid_dim=[7, 9, 7, 9], u_grid_dim=[3], xbar=[0.0, PI, 0.0, 0.0], INF=1000.0,
                            alpha=0.9, ctl=dict(K=[[0.3, -1.0, 0.3, -0.4]], ref=[0.0, PI, 0.0, 0.0]), snapshots=[1, 3, 10]),
}


def build_case(case, lookup=False):
    """Instantiate a case on the pyro_b200 mirrors -> (sys, grid_sys, cf)."""
    from pyro_b200 import costfunction, discretizer, systems
    sys_ = systems.SYSTEMS[case["system"]]()
    for key in ("x_lb", "x_ub", "u_lb", "u_ub"):
        if key in case:
            getattr(sys_, key)[:] = case[key]
    for key, val in case.get("sys_params", {}).items():
        setattr(sys_, key, val)
    grid = discretizer.GridDynamicSystem(sys_, case["x_grid_dim"], case["u_grid_dim"], case.get("dt", 0.05), lookup=lookup)
    cf = make_cost(costfunction, sys_, case)
    return sys_, grid, cf


def make_cost(costfunction, sys_, case):
    """The case's cost function on the given costfunction module (the mirrors' or the reference's) and system object."""
    kind = case.get("cost", "quadratic")
    if kind in ("quadratic", "domaincheck"):
        klass = costfunction.QuadraticCostFunction if kind == "quadratic" else costfunction.QuadraticCostFunctionWithDomainCheck
        cf = klass.from_sys(sys_)
        for key in ("Q", "R", "S"):
            if key in case:
                setattr(cf, key, np.diag(np.array(case[key], float)) if np.ndim(case[key]) == 1 else np.array(case[key], float))
    elif kind == "reach":
        cf = costfunction.Reachability(sys_.isavalidstate, np.array(case["xbar"], float))
    else:
        cf = costfunction.TimeCostFunction(np.array(case["xbar"], float))
    if "xbar" in case:
        cf.xbar = np.array(case["xbar"], float)
    for key in ("INF", "EPS"):
        if key in case:
            setattr(cf, key, case[key])
    return cf


def oracle_objects(case):
    """The same case on the independent NumPy oracle -> (GridOracle, cost)."""
    from oracle import np_oracle as npo
    spec = npo.SysSpec(case["system"], **case.get("sys_params", {}))
    for key in ("x_lb", "x_ub", "u_lb", "u_ub"):
        if key in case:
            setattr(spec, key, np.array(case[key], float))
    grid = npo.GridOracle(spec, case["x_grid_dim"], case["u_grid_dim"], case.get("dt", 0.05))
    if case.get("cost", "quadratic") in ("quadratic", "domaincheck"):   # on grid nodes the domain check never fires
        cost = npo.QuadCost(spec.n, spec.m)
        for key in ("Q", "R", "S"):
            if key in case:
                setattr(cost, key, np.diag(np.array(case[key], float)) if np.ndim(case[key]) == 1 else np.array(case[key], float))
    elif case["cost"] == "reach":
        cost = npo.ReachCost(case["xbar"])
    else:
        cost = npo.TimeCost(case["xbar"])
    if "xbar" in case:
        cost.xbar = np.array(case["xbar"], float)
    for key in ("INF", "EPS"):
        if key in case:
            setattr(cost, key, case[key])
    return grid, cost


def helicopter_tunnel_example(drone, costfunction, discretizer, x_grid_dim=(15, 13, 11), u_grid_dim=(5,)):
    """examples/demos_by_tool/dynamicprogramming/helicopter_tunnel.py of the reference on the given pyro modules (the real
    ones: the 3-D example systems live in pyro.dynamic, not in the mirrors), with a coarser grid: a 3-state plant whose
    isavalidstate holds obstacles, i.e. a table-mode problem.  -> (sys, grid_sys with look-up tables, cost function)."""
    sys_ = drone.ConstantSpeedHelicopterTunnel()
    sys_.obstacles = [[(2, 2), (4, 4)], [(8, 5), (10, 10)], [(14, 0), (16, 4)]]
    sys_.mass, sys_.vx, sys_.width = 0.1, 5.0, 1.0
    sys_.x_ub = np.array([+60, 10, +20])
    sys_.x_lb = np.array([-60, 0, +0])
    sys_.u_ub = np.array([+20])
    sys_.u_lb = np.array([-20])
    grid = discretizer.GridDynamicSystem(sys_, tuple(x_grid_dim), list(u_grid_dim), 0.05)
    qcf = costfunction.QuadraticCostFunctionWithDomainCheck.from_sys(sys_)
    qcf.xbar = np.array([0.0, 2.0, 20])
    qcf.INF, qcf.EPS = 100000, 0.2
    qcf.Q[0, 0], qcf.Q[1, 1], qcf.Q[2, 2] = 2.0, 200.0, 0.0
    qcf.R[0, 0] = 5.0
    qcf.S[0, 0], qcf.S[1, 1], qcf.S[2, 2] = 20.0, 50.0, 0.0
    return sys_, grid, qcf
