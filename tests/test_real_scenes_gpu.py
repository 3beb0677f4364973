"""GPU: the CUDA path on the REAL Argoverse-2 demo scene tensors (north_star: "outputs match the reference predictor on
identical Argoverse-2 scene tensors within 1e-3 relative, bit-exact for mode / branch index selection").
Golden vectors: tests/golden/real_demo_*.pt, dumped from the unmodified reference by oracle/make_golden_real.py."""
import copy

import numpy as np
import pytest
import torch

from conftest import rel_err
from test_real_scenes_cpu import DEMOS, compare_level_inputs, compare_tree, coord_ulp, load_real, net_inputs
from test_tree_oracle import TreeCfg

pytestmark = pytest.mark.gpu


def to_dev(x, dev):
    if torch.is_tensor(x):
        return x.to(dev)
    if isinstance(x, dict):
        return {k: to_dev(v, dev) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [to_dev(v, dev) for v in x]
    return x


def make_net(sd, dev, prec):
    from mind_b200.predictor import ScenePredNetB200
    net = ScenePredNetB200(None, dev)
    net.load_state_dict(sd)
    net.set_precision(prec)
    return net.to(dev).eval()


@pytest.mark.parametrize("prec,tol", [("fp32", 5e-5), ("f16tc", 5e-4)])
@pytest.mark.parametrize("name", DEMOS)
def test_forward_on_real_scene(ckpt_sd, name, prec, tol):
    dev = torch.device("cuda", 0)
    gold = load_real(name)
    net = make_net(ckpt_sd, dev, prec)
    cls, reg, aux = net(net.pre_process(to_dev(gold["data"], dev)))
    torch.cuda.synchronize()
    assert np.abs(cls[0].cpu().numpy() - gold["cls"][0].numpy()).max() < max(tol, 2e-6)
    assert (torch.argsort(-cls[0][0].cpu()) == torch.argsort(-gold["cls"][0][0])).all(), "mode order differs"
    e_reg, e_vel = rel_err(reg[0], gold["reg"][0]), rel_err(aux[0][0], gold["vel"][0])
    print("%s %s rel err reg %.3e vel %.3e" % (name, prec, e_reg, e_vel))
    assert e_reg < tol and e_vel < tol


def fragile_depths(ckpt_sd, gold, delta=0.05):
    """depth levels at which the reference algorithm itself has no well-defined node set on this scene: the oracle tree
    (CPU) reports, for every keep / merge decision of the greedy merge (:396-410), the distance of the deciding topology
    difference from the pi/6 threshold; a decision closer than `delta` rad (~0.2 m at 4 m range, the size of the
    reference's own fp32 noise at these levels, see compare_tree) can fall either way."""
    from oracle.tree_oracle import TreeOracle
    from test_tree_oracle import OracleNet
    t = TreeOracle(OracleNet(ckpt_sd), 50, 50, TreeCfg())
    t.reset()
    t.set_target_lane(gold["lane"], gold["info"])
    t.lane_graph = copy.deepcopy(gold["graph"])
    t.rollout(copy.deepcopy(gold["data"]))
    return sorted({d for d, _, m in t.merge_margins if abs(m) < delta}), min(abs(m) for _, _, m in t.merge_margins)


@pytest.mark.parametrize("prec", ["fp32", "f16tc"])
@pytest.mark.parametrize("name", DEMOS)
def test_tree_on_real_scene(ckpt_sd, name, prec):
    from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
    dev = torch.device("cuda", 0)
    gold = load_real(name)
    gen = ScenarioTreeGeneratorB200(dev, make_net(ckpt_sd, dev, prec), 50, 50, TreeCfg())
    gen.reset()
    gen.set_target_lane(gold["lane"], gold["info"])
    gen.lane_graph = copy.deepcopy(gold["graph"])
    trees = gen.rollout(copy.deepcopy(gold["data"]))
    flat = {k: (n.parent_key, float(n.data[0]), n.data[1], n.data[2], n.data[3]) for t in trees for k, n in t.nodes.items()}
    frag, closest = fragile_depths(ckpt_sd, gold)
    print("%s %s: %d nodes, levels %s, closest merge decision %.4f rad from the threshold, fragile depths %s" %
          (name, prec, len(flat), gen.net_batches, closest, frag))
    # every real scene has fewer than 128 tokens: in the tensor-core mode they take the exact tier, so both precisions
    # must reproduce the reference's branch / merge decisions node for node (north_star: bit-exact branch selection),
    # including demo_3's decision 0.0176 rad from its threshold
    compare_tree(flat, gen.net_batches, gold, 1e-3)


@pytest.mark.parametrize("name", DEMOS)
def test_tree_level_inputs_on_real_scene(ckpt_sd, name):
    """k_tree_update (window slide, re-normalisation, actor features, anchors, high-level command) vs the level inputs
    the reference built in update_obser, to a few ulp of the scene's global coordinates; the dense RPE the reference
    feeds is re-derived here from the anchors the kernel wrote (the device evaluates get_rpe inside k_edge_init)."""
    from mind_b200 import plumbing as P
    from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
    dev = torch.device("cuda", 0)
    gold = load_real(name)
    gen = ScenarioTreeGeneratorB200(dev, make_net(ckpt_sd, dev, "fp32"), 50, 50, TreeCfg())
    gen.graphs = False
    gen.reset()
    gen.set_target_lane(gold["lane"], gold["info"])
    gen.lane_graph = copy.deepcopy(gold["graph"])
    gen.rollout(copy.deepcopy(gold["data"]))
    torch.cuda.synchronize()
    assert len(gen._levels) == 1 + len(gold["level_inputs"])
    for lv, want in zip(gen._levels[1:], gold["level_inputs"]):
        actors, _, lanes, _, _, tgt_nodes, tgt_rpe = lv.net_in
        M = lv.geom[0].shape[0] // lv.F
        gc, gv = lv.geom[0].view(lv.F, M, 2).cpu(), lv.geom[1].view(lv.F, M, 2).cpu()
        got = dict(ACTORS=actors, LANES=lanes, TGT_NODES=tgt_nodes, TGT_RPE=tgt_rpe,
                   RPE=[P.pairwise_rpe(gc[f], gv[f]) for f in range(lv.F)])
        compare_level_inputs(got, want, coord_ulp(gold))


@pytest.mark.parametrize("prec,tol", [("fp32", 5e-5), ("f16tc", 5e-4)])
@pytest.mark.parametrize("name", DEMOS)
def test_forward_on_recorded_level_inputs(ckpt_sd, name, prec, tol):
    """the network on the level >= 1 batches exactly as the reference built them (3-4 scenes, lanes ~3-7 km away in the
    RPE because of the frame quirk of utils.py:171-177) vs the oracle on the same tensors"""
    from oracle.scene_pred_oracle import ScenePredOracle
    dev = torch.device("cuda", 0)
    gold = load_real(name)
    net = make_net(ckpt_sd, dev, prec)
    orc = ScenePredOracle(ckpt_sd)
    for li in gold["level_inputs"]:
        want = orc(net_inputs(li))
        cls, reg, aux = net(net.pre_process(to_dev(li, dev)))
        torch.cuda.synchronize()
        worst = 0.0
        for b in range(len(cls)):
            assert np.abs(cls[b].cpu().numpy() - want[0][b].numpy()).max() < max(tol, 2e-6)
            assert (torch.argsort(-cls[b][0].cpu()) == torch.argsort(-want[0][b][0])).all(), "mode order differs"
            e = max(rel_err(reg[b], want[1][b]), rel_err(aux[b][0], want[2][b][0]))
            worst = max(worst, e)
        print("%s %s recorded level batch of %d: worst rel err %.3e" % (name, prec, len(cls), worst))
        assert worst < tol
