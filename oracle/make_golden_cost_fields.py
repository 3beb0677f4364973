"""Golden cost fields from the UNMODIFIED reference TrajectoryTreeOptimizer (init_warm_start_cost_tree / init_cost_tree,
planners/mind/trajectory_tree.py:20-124) on the demo_2 scenario tree of tests/golden/real_demo_2.pt, on a small
non-square grid so that the fixture stays small.    python -m oracle.make_golden_cost_fields
Build container only (needs /root/reference; Theano through mind_b200.compat)."""
import os
import sys

import numpy as np
import torch

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRID, RES = (28, 20), 1.5


def scenario_trees(flat, tree_mod):
    """{key: (parent, prob, trajs, covs, tgt)} -> the reference's List[Tree] (one per depth-0 node)"""
    kids = {}
    for k, v in flat.items():
        kids.setdefault(v[0], []).append(k)
    trees = []
    for rk in sorted(k for k, v in flat.items() if v[0] is None):
        t = tree_mod.Tree()
        stack = [rk]
        while stack:
            k = stack.pop(0)
            parent, prob, trajs, covs, tgt = flat[k]
            t.add_node(tree_mod.Node(k, parent, [prob, trajs, covs, tgt]))
            stack += sorted(kids.get(k, []))
        trees.append(t)
    return trees


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
    state, ctrl = np.array([ego_xy[0], ego_xy[1], 6.0, 0.1]), np.array([0.2, 0.01])
    out = {"grid": np.array(GRID), "res": np.float64(RES), "state": state, "ctrl": ctrl, "lane": np.asarray(gold["lane"], dtype=np.float64)}
    for ti, st in enumerate(scenario_trees(gold["tree"], tree_mod)):
        for tag, init in (("warm", opt.init_warm_start_cost_tree), ("full", opt.init_cost_tree)):
            init(st, state, ctrl, out["lane"], 8.0)
            nodes = opt.cost_tree.tree.nodes
            keys = [k for k in nodes if k != -1]
            out["t%d/%s/fields" % (ti, tag)] = np.stack([nodes[k].data[0][0].cost_field for k in keys])
            out["t%d/%s/links" % (ti, tag)] = np.array([[k, nodes[k].parent_key] for k in keys])
            out["t%d/%s/offset" % (ti, tag)] = np.asarray(nodes[keys[0]].data[0][0].offset)
        print("tree", ti, "root", st.get_root().key, "fields", out["t%d/full/fields" % ti].shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cost_fields_demo_2.npz"), **out)


if __name__ == "__main__":
    main()
