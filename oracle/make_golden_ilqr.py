"""Golden iLQR solutions from the UNMODIFIED reference (TrajectoryTreeOptimizer: init_warm_start_cost_tree -> warm_start_solve
-> init_cost_tree -> solve, planners/mind/trajectory_tree.py + planners/ilqr/*) on the demo_2 scenario trees, on a 64 x 64
grid of 1.6 m cells so that the fixture stays small and the run short.        python -m oracle.make_golden_ilqr
Build container only (needs /root/reference; Theano through mind_b200.compat.theano_lite: exact Jacobians)."""
import os
import sys

import numpy as np
import torch

from oracle import ref_loader
from oracle.make_golden_cost_fields import scenario_trees

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRID, RES, TARGET_VEL = (64, 64), 1.6, 8.0


def main():
    from mind_b200 import compat
    compat.install()
    sys.path.insert(0, ref_loader.REF_ROOT)
    from planners.basic import tree as tree_mod
    from planners.mind.trajectory_tree import TrajectoryTreeOptimizer
    from planners.mind.configs.planning.demo_2 import TrajTreeCfg
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "real_demo_2.pt"), weights_only=False)
    cfg = TrajTreeCfg()
    for c in (cfg.w_opt_cfg, cfg.opt_cfg):
        c["smooth_grid_size"], c["smooth_grid_res"] = GRID, RES
    opt = TrajectoryTreeOptimizer(cfg)
    ego_xy = gold["tree"]["0_0_0"][2][0, 0]
    lane = np.asarray(gold["lane"], dtype=np.float64)
    d = lane[1:] - lane[:-1]
    k = int(np.argmin(np.linalg.norm(lane - ego_xy, axis=1)))
    heading = float(np.arctan2(d[min(k, len(d) - 1), 1], d[min(k, len(d) - 1), 0]))
    state, ctrl = np.array([ego_xy[0], ego_xy[1], 6.0, heading]), np.array([0.2, 0.01])
    out = {"grid": np.array(GRID), "res": np.float64(RES), "state": state, "ctrl": ctrl, "lane": lane, "target_vel": np.float64(TARGET_VEL),
           "dt": np.float64(cfg.dt)}
    for ti, st in enumerate(scenario_trees(gold["tree"], tree_mod)):
        opt.init_warm_start_cost_tree(st, state, ctrl, lane, TARGET_VEL)
        xs_w, us_w = opt.warm_start_solve()
        out["t%d/warm/xs" % ti], out["t%d/warm/us" % ti] = xs_w.copy(), us_w.copy()
        opt.init_cost_tree(st, state, ctrl, lane, TARGET_VEL)
        tt = opt.solve(us_w)
        out["t%d/full/xs" % ti], out["t%d/full/us" % ti] = opt.ilqr.xs.copy(), opt.ilqr.us.copy()
        print("tree", ti, "nodes", len(xs_w), "warm J %.4f" % opt.ilqr.J_opt, "final ego", np.round(opt.ilqr.xs[-1][:4], 3), "tree nodes", len(tt.nodes))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ilqr_demo_2.npz"), **out)


if __name__ == "__main__":
    main()
