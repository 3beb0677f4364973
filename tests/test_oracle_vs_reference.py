"""CPU, build container only: oracle restatement vs the UNMODIFIED reference modules imported
from /root/reference through the shims (skipped where the reference tree is absent)."""
import pytest
import torch

from conftest import rel_err
from mind_b200 import synth
from oracle import ref_loader
from oracle import scene_pred_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref(ckpt_sd):
    net, _ = ref_loader.build_reference_net(ckpt_sd)
    return net


def test_stages_and_outputs(ref, ckpt_sd):
    data = synth.batch_ragged(batch=2, seed=3, na_rng=(2, 6), nl_rng=(5, 12), seed0=77)
    orc = O.ScenePredOracle(ckpt_sd)
    st = orc.stages(data)
    with torch.no_grad():
        a = ref.actor_net(data[0]); l = ref.lane_net(data[2])
        a2, l2, c2 = ref.fusion_net(a, data[1], l, data[3], data[4])
        rc, rr, ra = ref(data)
    assert rel_err(st["actor_feat"], a) < 1e-5
    assert rel_err(st["lane_feat"], l) < 1e-5
    assert rel_err(st["actors"], a2) < 2e-5 and rel_err(st["cls"], c2) < 2e-5
    oc, orr, oa = orc(data)
    for b in range(2):
        assert (oc[b] - rc[b]).abs().max() < 1e-6
        assert rel_err(orr[b], rr[b]) < 2e-5 and rel_err(oa[b][0], ra[b][0]) < 2e-5
        assert tuple(orr[b].shape) == tuple(rr[b].shape) and tuple(oa[b][2].shape) == tuple(ra[b][2].shape)


def test_single_layer_matches_module(ref):
    """RelaFusionLayer index convention (query j attends keys i over memory[i, j])."""
    torch.manual_seed(0)
    n = 9
    node, edge = torch.randn(n, 128), torch.randn(n, n, 128)
    for li in (0, 5):
        mod = ref.fusion_net.fuse_scene.fusion[li]
        with torch.no_grad():
            x_ref, e_ref, _ = mod(node, edge, None)
        p = O.Params(ref.state_dict(), "fusion_net.fuse_scene.fusion.%d." % li)
        x, e = O.rela_fusion_layer(node, edge, p, update_edge=(li != 5))
        assert rel_err(x, x_ref) < 1e-5 and rel_err(e, e_ref) < 1e-5


def test_get_rpe_matches(ref):
    ns = ref_loader.load()
    s = synth.scene_s1(11, 5, 9, with_geom=True)
    r, _ = ns.utils.get_rpe(s["ctrs"], s["vecs"])
    assert torch.equal(r, O.get_rpe(s["ctrs"], s["vecs"])) and torch.equal(r, s["rpe"])


# ---- AIME tree: oracle restatement vs the live reference ScenarioTreeGenerator --------------------
@pytest.mark.parametrize("v", [0, 1, 2])
def test_tree_oracle_matches_reference_generator(ref, v):
    import copy
    import numpy as np
    from oracle.make_golden_tree import VARIANTS, reference_tree
    from oracle.tree_oracle import TreeOracle, flatten_trees
    ns = ref_loader.load()
    want, levels = reference_tree(ns, ref, VARIANTS[v])
    t = TreeOracle(ref, 50, 50, ns.plan_cfg.ScenTreeCfg())
    data, lane, info, graph = synth.scene_s3(**VARIANTS[v])
    t.reset()
    t.set_target_lane(lane, info)
    t.lane_graph = copy.deepcopy(graph)
    got = flatten_trees(t.rollout(data))
    assert t.net_batches == levels and sorted(got) == sorted(want)
    for k in want:
        assert got[k][0] == want[k][0]
        assert abs(got[k][1] - want[k][1]) < 1e-6
        assert np.abs(got[k][2] - want[k][2]).max() < 1e-4 and np.abs(got[k][3] - want[k][3]).max() < 1e-5
        assert np.array_equal(got[k][4], want[k][4])


# ---- host plumbing mirrors (mind_b200/plumbing.py) vs the reference utils ------------------------
def test_collate_and_frames_match_reference(ref):
    import copy
    from mind_b200 import plumbing as P
    ns = ref_loader.load()
    data, lane, info, graph = synth.scene_s3()
    # un-collate: rebuild the per-scene dict and run both collates
    one = {k: (v[0] if isinstance(v, list) else v) for k, v in data.items()
           if k in ("ORIG", "ROT", "TRAJS", "LANE_GRAPH", "TGT_PTS", "TGT_ANCH", "RPE")}
    one["TGT_NODES"] = data["TGT_NODES"][0]
    one["TGT_RPE"] = data["TGT_RPE"].view(1, 5, 2, 2)[0]
    a = ns.utils.collate_fn([copy.deepcopy(one), copy.deepcopy(one)])
    b = P.collate_scenes([copy.deepcopy(one), copy.deepcopy(one)])
    for key in ("ACTORS", "LANES", "TGT_NODES", "TGT_RPE"):
        assert torch.equal(a[key].float(), b[key].float()), key
    assert all(torch.equal(x, y) for x, y in zip(a["ACTOR_IDCS"], b["ACTOR_IDCS"]))
    assert all(torch.equal(x, y) for x, y in zip(a["LANE_IDCS"], b["LANE_IDCS"]))
    # frames and high-level command
    pos, ang = torch.randn(50, 2), torch.randn(50)
    o1, r1, t1 = ns.utils.get_origin_rotation(pos, ang, torch.device("cpu"))
    o2, r2, t2 = P.origin_rotation(pos, ang)
    assert torch.equal(o1, o2) and torch.allclose(r1, r2) and torch.equal(t1, t2)
    gen = ns.scenario_tree.ScenarioTreeGenerator(torch.device("cpu"), ref, 50, 50, ns.plan_cfg.ScenTreeCfg())
    gen.set_target_lane(lane, info)
    orig, rot = torch.tensor([40.0, 1.0]), torch.tensor([[0.9, -0.43588989], [0.43588989, 0.9]])
    p1, n1, a1 = gen.get_high_level_command(orig, rot, 7.0)
    p2, n2, a2 = P.high_level_command(gen.target_lane, gen.target_lane_info, orig, rot, 7.0, 5.0)
    assert torch.equal(p1, p2) and torch.allclose(n1.float(), n2.float(), atol=1e-6)
    assert torch.allclose(a1[0], a2[0], atol=1e-6) and torch.allclose(a1[1], a2[1], atol=1e-6)
