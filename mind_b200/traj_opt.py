"""Trajectory-tree optimisation step of MIND on the accelerated pieces of this repo:
cost fields from the CUDA kernels (mind_b200.cost_field, or any fields handed in) + the native tree iLQR
(`mind_ilqr_tree_solve`, csrc/ilqr_tree.cpp).

`solve_tree(...)` is the functional core: one warm-start or full solve on one scenario tree, mirroring
`TrajectoryTreeOptimizer.init_*cost_tree` + `ilqr.fit` (planners/mind/trajectory_tree.py:20-147).
`TrajectoryTreeOptimizerB200` keeps the reference class's call surface (init_warm_start_cost_tree / warm_start_solve /
init_cost_tree / solve, planner.py:171-175) so `MINDPlanner.get_traj_tree` can use it unchanged.
"""
import ctypes as C

import numpy as np

from . import cost_field as CF
from . import lib as _lib

WHEELBASE = 2.5          # trajectory_tree.py:16 (`_get_dynamic_model(self.config.dt, 2.5)`)


def ilqr_solve(parent, x0, dt, offset, xs_grid, ys_grid, res, fields, w_state, des_state, w_con, lower, upper, w_ctrl, us_init,
               max_iter=100, wheelbase=WHEELBASE):
    """Thin binding of mind_ilqr_tree_solve (host code).  Arrays are converted to contiguous fp64 / int32.
    Returns (xs [n,6], us [n,2], iterations, cost)."""
    L = _lib.load()
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    n = len(parent)
    arrs = dict(x0=f64(x0), off=f64(offset), xs=f64(xs_grid), ys=f64(ys_grid), fields=f64(fields), ws=f64(w_state), des=f64(des_state),
                wc=f64(w_con), lo=f64(lower), up=f64(upper), wu=f64(w_ctrl), us=f64(us_init))
    assert arrs["fields"].shape == (n, len(arrs["ys"]), len(arrs["xs"])), arrs["fields"].shape
    assert arrs["ws"].shape == (n, 6, 6) and arrs["wc"].shape == (n, 6, 6) and arrs["wu"].shape == (n, 2, 2)
    assert arrs["des"].shape == (n, 6) and arrs["us"].shape == (n, 2)
    xs_out, us_out = np.empty((n, 6)), np.empty((n, 2))
    it, cost = C.c_int32(0), C.c_double(0.0)
    p = _lib.MindIlqrTree()
    ptr = lambda a: a.ctypes.data
    p.n_nodes, p.parent, p.x0, p.dt, p.wheelbase = n, ptr(parent), ptr(arrs["x0"]), float(dt), float(wheelbase)
    p.gx, p.gy, p.res, p.field_offset = len(arrs["xs"]), len(arrs["ys"]), float(res), ptr(arrs["off"])
    p.xs_grid, p.ys_grid, p.fields = ptr(arrs["xs"]), ptr(arrs["ys"]), ptr(arrs["fields"])
    p.w_state, p.des_state, p.w_con, p.lower, p.upper, p.w_ctrl = (ptr(arrs["ws"]), ptr(arrs["des"]), ptr(arrs["wc"]), ptr(arrs["lo"]),
                                                                    ptr(arrs["up"]), ptr(arrs["wu"]))
    p.max_iter, p.us_init, p.xs_out, p.us_out = int(max_iter), ptr(arrs["us"]), ptr(xs_out), ptr(us_out)
    p.iterations, p.cost = C.addressof(it), C.addressof(cost)
    if L.mind_ilqr_tree_solve(C.byref(p)) != 0:
        raise RuntimeError(L.mind_ilqr_last_error().decode())
    return xs_out, us_out, int(it.value), float(cost.value)


def node_weights(probs, cfg, target_vel):
    """per-node potentials of trajectory_tree.py:40-47 / :109-117: weights scale with the scenario probability"""
    probs = np.asarray([float(p) for p in probs])
    n = len(probs)
    w_state = probs[:, None, None] * np.asarray(cfg["w_des_state"], dtype=np.float64)[None]
    w_con = probs[:, None, None] * np.asarray(cfg["w_state_con"], dtype=np.float64)[None]
    w_ctrl = probs[:, None, None] * np.asarray(cfg["w_ctrl"], dtype=np.float64)[None]
    des = np.tile(np.array([0, 0, target_vel, 0.0, 0.0, 0.0], dtype=np.float64), (n, 1))
    return w_state, des, w_con, w_ctrl


def solve_tree(scen_tree, x0, target_lane, target_vel, cfg, dt, us_init=None, warm=False, device="cuda", fields=None):
    """One solve on one scenario tree.  `fields` = dict from cost_field.cost_fields (computed here on `device` when None).
    Returns (xs, us, info) with info = dict(links, probs, iterations, cost, offset, xx, yy)."""
    out = fields if fields is not None else CF.cost_fields(scen_tree, x0, target_lane, cfg, device, warm=warm)
    links = out["links"]
    parent = [last for _, last in links]
    n = len(parent)
    w_state, des, w_con, w_ctrl = node_weights(out["probs"], cfg, target_vel)
    if us_init is None:
        us_init = np.zeros((n, 2))
    xs, us, it, cost = ilqr_solve(parent, x0, dt, out["offset"], out["xx"][0], out["yy"][:, 0], cfg["smooth_grid_res"], out["fields"],
                                  w_state, des, w_con, cfg["state_lower_bound"], cfg["state_upper_bound"], w_ctrl, us_init)
    return xs, us, dict(out, iterations=it, cost=cost)


class TrajectoryTreeOptimizerB200:
    """Call surface of planners/mind/trajectory_tree.py::TrajectoryTreeOptimizer (planner.py:171-175)."""

    def __init__(self, config, device="cuda"):
        self.config, self.device = config, device
        self.debug = None
        self._job = None

    def _get_init_state(self, init_state, init_ctrl):                        # :150-152
        return np.array([init_state[0], init_state[1], init_state[2], init_state[3], init_ctrl[0], init_ctrl[1]])

    def init_warm_start_cost_tree(self, scen_tree, init_state, init_ctrl, target_lane, target_vel):
        self._job = (scen_tree, self._get_init_state(init_state, init_ctrl), target_lane, target_vel, self.config.w_opt_cfg, True)

    def init_cost_tree(self, scen_tree, init_state, init_ctrl, target_lane, target_vel):
        self._job = (scen_tree, self._get_init_state(init_state, init_ctrl), target_lane, target_vel, self.config.opt_cfg, False)

    def _run(self, us_init):
        scen_tree, x0, lane, vel, cfg, warm = self._job
        return solve_tree(scen_tree, x0, lane, vel, cfg, self.config.dt, us_init=us_init, warm=warm, device=self.device)

    def warm_start_solve(self, us_init=None):                                 # :125-130
        xs, us, _ = self._run(us_init)
        return xs, us

    def solve(self, us_init=None):                                            # :132-147
        try:
            from planners.basic.tree import Tree, Node
        except ImportError:
            from .scenario_tree import Tree, Node
        xs, us, info = self._run(us_init)
        x0 = self._job[1]
        tree = Tree()
        tree.add_node(Node(-1, None, [x0, np.zeros(self.config.action_size)]))
        for idx, last in info["links"]:
            tree.add_node(Node(idx, last, [xs[idx], us[idx]]))
        return tree
