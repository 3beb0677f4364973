"""Golden AIME trees from the UNMODIFIED reference (ScenarioTreeGenerator + ScenePredNet, shipped
checkpoint) on the kinematic S3 scenes of mind_b200/synth.py.   python -m oracle.make_golden_tree
Writes tests/golden/tree_s3.npz: per variant v and node key k:  v{v}/{k}/parent|prob|trajs|covs|tgt,
plus v{v}/levels (net batch size per depth level)."""
import copy
import os

import numpy as np
import torch

from oracle import ref_loader
from oracle.tree_oracle import flatten_trees
from mind_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tree_s3.npz")

VARIANTS = [dict(),
            dict(x0=(50, 58, 44, 70, 62, 40, 90), y0=(0, 3.5, -3.5, 0, 3.5, 0, -3.5), v=(8, 4, 9, 3, 7, 10, 2)),
            dict(x0=(100, 108, 92, 120, 112, 96, 130, 85), y0=(0, 3.5, -3.5, 0, 3.5, 3.5, -3.5, 0), v=(5, 9, 3, 8, 2, 10, 6, 12))]


def reference_tree(ns, net, args):
    gen = ns.scenario_tree.ScenarioTreeGenerator(torch.device("cpu"), net, 50, 50, ns.plan_cfg.ScenTreeCfg())
    data, lane, info, graph = synth.scene_s3(**args)
    gen.reset()
    gen.set_target_lane(lane, info)
    gen.lane_graph = copy.deepcopy(graph)
    gen.init_scenario_tree(data)                      # scenario_tree.py:41 (process_data replaced by scene_s3)
    levels = [1]
    nodes = gen.get_branch_set()
    while nodes:                                      # scenario_tree.py:44-55
        batch = ns.utils.collate_fn([n.data.obs_data for n in nodes])
        levels.append(len(nodes))
        gen.create_nodes(gen.prune_merge(batch, gen.predict_scenes(batch)))
        gen.decide_branch()
        nodes = gen.get_branch_set()
    return flatten_trees(gen.get_scenario_tree()), levels


def main():
    ns = ref_loader.load()
    net, _ = ref_loader.build_reference_net()
    out = {}
    for v, args in enumerate(VARIANTS):
        flat, levels = reference_tree(ns, net, args)
        out["v%d/levels" % v] = np.array(levels)
        out["v%d/keys" % v] = np.array(sorted(flat.keys()))
        for k, (parent, prob, trajs, covs, tgt) in flat.items():
            out["v%d/%s/parent" % (v, k)] = np.array(parent if parent is not None else "")
            out["v%d/%s/prob" % (v, k)] = np.float32(prob)
            out["v%d/%s/trajs" % (v, k)] = trajs
            out["v%d/%s/covs" % (v, k)] = covs
            out["v%d/%s/tgt" % (v, k)] = tgt
        print("variant", v, "levels", levels, "nodes", sorted(flat.keys()))
    np.savez_compressed(OUT, **out)


if __name__ == "__main__":
    main()
