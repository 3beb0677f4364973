"""CPU: the oracle on the REAL Argoverse-2 demo scenes (BASELINE.json configs 1 and 3).

tests/golden/real_demo_*.pt were dumped by oracle/make_golden_real.py from the UNMODIFIED reference (its own loader,
agents, process_data, ScenePredNet and branch_aime; av2 / shapely through mind_b200.compat).  Here the oracle's network
and tree restatements are pinned on those scenes: outputs within 5e-5 (fp32 summation order), tree node ids / parents / level sizes exact."""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from test_tree_oracle import OracleNet, TreeCfg

DEMOS = ["demo_1", "demo_2", "demo_3", "demo_4"]
# the closed-loop agent's first plan call (sim time 4.0 s): 41 observed steps, every track front-padded, 6 first-level nodes
FIRST_PLAN = ["demo_3_t4"]


def load_real(name):
    return torch.load(os.path.join(GOLDEN, "real_%s.pt" % name), weights_only=False)


def net_inputs(data):
    return tuple(data[k] for k in ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"])


def compare_tree(flat, levels, gold, tol):
    """Structure (node ids, parents, level batch sizes = every mode / branch index decision) must be identical.
    Numbers: nodes of the first level come from the root forward on identical inputs -> `tol`.  Deeper nodes are limited
    by the REFERENCE's own conditioning, not by the implementation under test: it carries global-frame coordinates of
    3-7 km in fp32 (1 ulp = 2.4e-4 .. 4.9e-4 m) and re-derives 0.05-1 m displacements from them (scenario_tree.py:341,
    498-511), so two fp32 evaluations of the same formulas already differ by ~1e-2 in probability there (measured:
    oracle vs reference 6.6e-3, covariances 4.6e-2).  They are held to 3e-2 / 0.1 / 0.3 m; the inputs of those levels
    are checked separately at the ulp level (compare_level_inputs)."""
    assert list(levels) == list(gold["levels"])
    want = gold["tree"]
    assert sorted(flat) == sorted(want)
    for k, (parent, prob, trajs, covs, tgt) in want.items():
        g = flat[k]
        assert (g[0] or "") == (parent or "")
        deep = not k.startswith("0_")
        assert abs(g[1] - prob) < (3e-2 if deep else tol), (k, g[1], prob)
        assert g[2].shape == trajs.shape and g[3].shape == covs.shape
        if deep:
            assert np.abs(g[2] - trajs).max() < 0.3 and np.abs(g[3] - covs).max() < 0.1 * max(1.0, np.abs(covs).max()), k
        else:
            assert np.abs(g[2] - trajs).max() < tol * max(1.0, np.abs(trajs).max()), k
            assert np.abs(g[3] - covs).max() < tol * max(1.0, np.abs(covs).max()), k
        assert np.abs(np.asarray(g[4]) - np.asarray(tgt)).max() < 1e-4 * max(1.0, np.abs(np.asarray(tgt)).max())


def oracle_margins(ckpt_sd, gold):
    """merge decisions of the oracle tree on this scene: [(depth, scene, margin in rad from the pi/6 threshold)]"""
    from oracle.tree_oracle import TreeOracle
    t = TreeOracle(OracleNet(ckpt_sd), 50, 50, TreeCfg())
    t.reset()
    t.set_target_lane(gold["lane"], gold["info"])
    t.lane_graph = copy.deepcopy(gold["graph"])
    t.rollout(copy.deepcopy(gold["data"]))
    return t.merge_margins


def compare_tree_with_margins(flat, levels, gold, tol, margins, delta=0.15):
    """compare_tree, except that a node may be present on one side only when its parent scene has a greedy-merge decision
    (:396-410) within `delta` rad of the pi/6 threshold.  The topology signature is the angle swept by (exo - ego); at
    depth >= 1 the reference's own fp32 noise on positions is ~0.1 m (see compare_tree), i.e. ~0.1 rad for an actor 1-2 m
    from the ego, so such a decision has no well-defined outcome.  A flip is only tolerated where it does not renumber
    the next frontier (equal level sizes); returns the tolerated keys."""
    fragile = {(d, b) for d, b, m in margins if abs(m) < delta}
    want = gold["tree"]
    odd = sorted(set(flat) ^ set(want))
    if not odd:
        compare_tree(flat, levels, gold, tol)
        return []
    assert list(levels) == list(gold["levels"]), "a flipped decision changed the frontier sizes"
    for k in odd:
        d, b, _ = (int(x) for x in k.split("_"))
        assert (d, b) in fragile, "node %s differs although no merge decision of scene (%d, %d) is near its threshold" % (k, d, b)
    common = {k: v for k, v in want.items() if k in flat}
    # sibling probabilities are renormalised over the kept siblings (:240-262): a flip changes them for that family
    fam = {want.get(k, flat.get(k))[0] for k in odd}
    fam |= {flat[k][0] for k in odd if k in flat}
    keep = {k: v for k, v in common.items() if v[0] not in fam}
    compare_tree({k: flat[k] for k in keep}, levels, dict(gold, tree=keep), tol)
    return odd


def coord_ulp(gold):
    """fp32 spacing at the scene's global coordinates (the reference's own resolution limit for level >= 1 inputs)"""
    return float(np.spacing(np.float32(np.abs(gold["data"]["ORIG"][0].numpy()).max())))


def compare_level_inputs(got, want, ulp):
    """got / want: dicts with ACTORS [A,14,48], LANES, RPE (list of [5,M,M]), TGT_NODES, TGT_RPE of one depth level."""
    a, b = got["ACTORS"].float().cpu(), want["ACTORS"].float()
    assert a.shape == b.shape
    d = (a - b).abs().amax(dim=(0, 2))
    assert d[0:2].max() < 8 * ulp, ("displacement", d[0:2], ulp)          # differences of global fp32 coordinates
    assert d[2:6].max() < 2e-4, ("heading / velocity", d[2:6])
    assert d[6:14].max() == 0, ("type one-hot / pad flag", d[6:14])
    assert torch.equal(got["LANES"].float().cpu(), want["LANES"].float())
    assert (got["TGT_NODES"].float().cpu() - want["TGT_NODES"].float()).abs().max() < 1e-4
    assert (got["TGT_RPE"].float().cpu() - want["TGT_RPE"].float()).abs().max() < 1e-4
    for r1, r2 in zip(got["RPE"], want["RPE"]):
        r1 = r1["scene"] if isinstance(r1, dict) else r1
        r2 = r2["scene"] if isinstance(r2, dict) else r2
        assert (r1.float().cpu() - r2.float()).abs().max() < 2e-3, "RPE"   # values up to ~70 (root-frame lanes, utils.py:171-177)


@pytest.mark.parametrize("name", DEMOS + FIRST_PLAN)
def test_oracle_forward_on_real_scene(ckpt_sd, name):
    from oracle.scene_pred_oracle import ScenePredOracle
    gold = load_real(name)
    cls, reg, aux = ScenePredOracle(ckpt_sd)(net_inputs(gold["data"]))
    assert len(cls) == 1
    assert np.abs(cls[0].numpy() - gold["cls"][0].numpy()).max() < 2e-6
    assert (torch.argsort(-cls[0][0]) == torch.argsort(-gold["cls"][0][0])).all()
    assert rel_err(reg[0], gold["reg"][0]) < 5e-5 and rel_err(aux[0][0], gold["vel"][0]) < 5e-5


@pytest.mark.parametrize("name", DEMOS + FIRST_PLAN)
def test_oracle_tree_on_real_scene(ckpt_sd, name):
    from oracle.tree_oracle import TreeOracle, flatten_trees
    gold = load_real(name)
    t = TreeOracle(OracleNet(ckpt_sd), 50, 50, TreeCfg())
    t.reset()
    t.set_target_lane(gold["lane"], gold["info"])
    t.lane_graph = copy.deepcopy(gold["graph"])
    flat = flatten_trees(t.rollout(copy.deepcopy(gold["data"])))
    flips = compare_tree_with_margins(flat, t.net_batches, gold, 1e-4, t.merge_margins)
    assert flips == ([] if name in DEMOS else ["1_0_3"])      # first-plan scene: one decision 0.06 rad from its threshold


@pytest.mark.parametrize("name", DEMOS + FIRST_PLAN)
def test_oracle_level_inputs_on_real_scene(ckpt_sd, name):
    """update_obser (:467-567) of the oracle tree vs the level inputs the reference built, to a few ulp of the coordinates"""
    from oracle.tree_oracle import TreeOracle
    gold = load_real(name)
    net = OracleNet(ckpt_sd)
    seen = []
    pre = net.pre_process
    net.pre_process = lambda d: (seen.append({k: copy.deepcopy(d[k]) for k in ("ACTORS", "LANES", "RPE", "TGT_NODES", "TGT_RPE")}), pre(d))[1]
    t = TreeOracle(net, 50, 50, TreeCfg())
    t.reset()
    t.set_target_lane(gold["lane"], gold["info"])
    t.lane_graph = copy.deepcopy(gold["graph"])
    t.rollout(copy.deepcopy(gold["data"]))
    assert len(seen) == 1 + len(gold["level_inputs"])
    for got, want in zip(seen[1:], gold["level_inputs"]):
        compare_level_inputs(got, want, coord_ulp(gold))
