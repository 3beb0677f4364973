"""GPU: AIME scenario tree on the B200 path (CUDA tree-step kernels + CUDA predictor) vs the golden
trees dumped from the unmodified reference: node ids / parents / per-level batch sizes exact
(bit-exact mode and branch index selection), probabilities, trajectories, covariances within 1e-3."""
import copy

import pytest
import torch

from conftest import load_golden
from test_tree_oracle import TreeCfg, compare_with_golden

pytestmark = pytest.mark.gpu


def flatten(trees):
    out = {}
    for t in trees:
        for k, n in t.nodes.items():
            out[k] = (n.parent_key, float(n.data[0]), n.data[1], n.data[2], n.data[3])
    return out


@pytest.mark.parametrize("prec", ["fp32", "f16tc"])
@pytest.mark.parametrize("v", [0, 1, 2])
def test_tree_vs_golden(ckpt_sd, v, prec):
    from mind_b200 import synth
    from mind_b200.predictor import ScenePredNetB200
    from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
    from oracle.make_golden_tree import VARIANTS
    dev = torch.device("cuda", 0)
    net = ScenePredNetB200(None, dev)
    net.load_state_dict(ckpt_sd)
    net.set_precision(prec)
    gen = ScenarioTreeGeneratorB200(dev, net, 50, 50, TreeCfg())
    data, lane, info, graph = synth.scene_s3(**VARIANTS[v])
    gen.reset()
    gen.set_target_lane(lane, info)
    gen.lane_graph = copy.deepcopy(graph)
    trees = gen.rollout(data)
    flat = flatten(trees)
    gold = load_golden("tree_s3.npz")
    compare_with_golden(flat, gen.net_batches, gold, v)
    for t in trees:
        for n in t.nodes.values():
            assert n.data[1].dtype.name == "float32" and n.data[1].ndim == 3 and n.data[2].shape[-1] == 1


def test_tree_vs_oracle_tree_new_scene(ckpt_sd):
    """a 10-actor scene that is not among the goldens: CUDA tree (exact fp32 predictor) vs the CPU oracle tree."""
    from mind_b200 import synth
    from mind_b200.predictor import ScenePredNetB200
    from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
    from oracle.tree_oracle import TreeOracle, flatten_trees
    from test_tree_oracle import OracleNet
    import numpy as np
    scene = dict(x0=(70, 78, 62, 90, 84, 66, 100, 55, 110, 48), y0=(0, 3.5, -3.5, 0, 3.5, 3.5, -3.5, 0, 3.5, -3.5),
                 v=(7, 3, 11, 5, 9, 2, 6, 12, 4, 8))
    dev = torch.device("cuda", 0)
    net = ScenePredNetB200(None, dev)
    net.load_state_dict(ckpt_sd)
    net.set_precision("fp32")
    gen = ScenarioTreeGeneratorB200(dev, net, 50, 50, TreeCfg())
    data, lane, info, graph = synth.scene_s3(**scene)
    gen.reset(); gen.set_target_lane(lane, info); gen.lane_graph = copy.deepcopy(graph)
    got = flatten(gen.rollout(data))
    orc = TreeOracle(OracleNet(ckpt_sd), 50, 50, TreeCfg())
    data, lane, info, graph = synth.scene_s3(**scene)
    orc.reset(); orc.set_target_lane(lane, info); orc.lane_graph = copy.deepcopy(graph)
    want = flatten_trees(orc.rollout(data))
    assert gen.net_batches == orc.net_batches
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k][0] == want[k][0] and abs(got[k][1] - want[k][1]) < 1e-3
        assert got[k][2].shape == want[k][2].shape and np.abs(got[k][2] - want[k][2]).max() < 1e-3 * max(1.0, np.abs(want[k][2]).max())
        assert np.abs(got[k][3] - want[k][3]).max() < 1e-3 * max(1.0, np.abs(want[k][3]).max())
