"""CPU: the oracle tree (oracle net + restated AIME generator) reproduces the golden trees dumped
from the unmodified reference (tests/golden/tree_s3.npz): identical node ids / parents / level
batch sizes (bit-exact index decisions), probabilities and trajectories within 1e-4."""
import copy

import numpy as np
import pytest
import torch

from conftest import load_golden
from mind_b200 import synth
from oracle.make_golden_tree import VARIANTS
from oracle.scene_pred_oracle import ScenePredOracle
from oracle.tree_oracle import TreeOracle, flatten_trees


class TreeCfg:      # planners/mind/configs/planning/demo_1.py:3-10 (ScenTreeCfg)
    max_depth = 5
    tar_dist_thres = 10.0
    tar_time_ahead = 5.0


class OracleNet:
    def __init__(self, sd):
        self.o = ScenePredOracle(sd)

    def pre_process(self, d):
        return tuple(d[k] for k in ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"])

    def __call__(self, x):
        return self.o(x)


def compare_with_golden(flat, levels, gold, v, tol=1e-3):
    assert list(gold["v%d/levels" % v]) == list(levels)
    assert sorted(flat.keys()) == list(gold["v%d/keys" % v])
    for k, (parent, prob, trajs, covs, tgt) in flat.items():
        assert (parent or "") == str(gold["v%d/%s/parent" % (v, k)])
        assert abs(prob - float(gold["v%d/%s/prob" % (v, k)])) < tol
        g = gold["v%d/%s/trajs" % (v, k)]
        assert trajs.shape == g.shape and np.abs(trajs - g).max() < tol * max(1.0, np.abs(g).max())
        g = gold["v%d/%s/covs" % (v, k)]
        assert covs.shape == g.shape and np.abs(covs - g).max() < tol * max(1.0, np.abs(g).max())
        assert np.abs(tgt - gold["v%d/%s/tgt" % (v, k)]).max() < 1e-4


@pytest.mark.parametrize("v", range(len(VARIANTS)))
def test_oracle_tree_vs_golden(ckpt_sd, v):
    gold = load_golden("tree_s3.npz")
    t = TreeOracle(OracleNet(ckpt_sd), 50, 50, TreeCfg())
    data, lane, info, graph = synth.scene_s3(**VARIANTS[v])
    t.reset()
    t.set_target_lane(lane, info)
    t.lane_graph = copy.deepcopy(graph)
    flat = flatten_trees(t.rollout(data))
    compare_with_golden(flat, t.net_batches, gold, v)
