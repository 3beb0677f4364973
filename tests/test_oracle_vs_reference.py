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
