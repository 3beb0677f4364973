"""GPU: scene sizes that make the fused layer's last query block a single-query item (N mod 16 == 1) with one, two and
three 128-key tiles, alone and mixed with ordinary scenes in one batch, vs the oracle.  (The 161-token benchmark scenes
cover the two-tile case in test_forward_gpu.py; this file adds the other shapes of tc_build_schedule's mode-1 items.)

In the tensor-core mode scenes below option "tc_min_tokens" (default 128) take the exact tier, so the default-mode cases
here mix both tiers in one batch; the `fused_only` cases force every scene through the fp16-operand fused kernel
(tc_min_tokens = 0) to keep its one-tile single-query path covered -- there the bound is the kernel's documented
small-scene error (operand rounding is not averaged out over 17 keys), not the parity tolerance."""
import pytest
import torch

from conftest import rel_err
from test_forward_gpu import TOL_FP32, TOL_TC, make_net, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,tol", [("fp32", TOL_FP32), ("f16tc", TOL_TC)])
@pytest.mark.parametrize("sizes", [[(4, 12)], [(16, 96)], [(32, 96)], [(8, 24), (32, 128), (5, 20), (16, 96)], [(64, 192)]])
def test_single_query_item_shapes(ckpt_sd, sizes, prec, tol):
    from mind_b200 import synth
    from oracle.scene_pred_oracle import ScenePredOracle
    dev = torch.device("cuda", 0)
    assert any((na + nl + 1) % 16 == 1 for na, nl in sizes)
    scenes = [synth.scene_s1(700 + 13 * i + na, na, nl) for i, (na, nl) in enumerate(sizes)]
    data = synth.batch_from_scenes(scenes)
    oc, orr, oa = ScenePredOracle(ckpt_sd)(data)
    net = make_net(ckpt_sd, dev, prec)
    cls, reg, aux = net(to_dev(data, dev))
    torch.cuda.synchronize()
    net.sync_check()
    worst = 0.0
    for b in range(len(sizes)):
        assert (cls[b].cpu() - oc[b]).abs().max() < max(tol, 1e-5)
        assert torch.equal(cls[b][0].cpu().argsort(descending=True), oc[b][0].argsort(descending=True))
        worst = max(worst, rel_err(reg[b], orr[b]), rel_err(aux[b][0], oa[b][0]))
    print("sizes %s %s: worst rel err %.3e" % ([na + nl + 1 for na, nl in sizes], prec, worst))
    assert worst < tol


@pytest.mark.parametrize("sizes", [[(4, 12)], [(16, 96)], [(8, 24), (32, 128), (5, 20), (16, 96)]])
def test_single_query_item_shapes_fused_only(ckpt_sd, sizes):
    from mind_b200 import synth
    from oracle.scene_pred_oracle import ScenePredOracle
    dev = torch.device("cuda", 0)
    scenes = [synth.scene_s1(700 + 13 * i + na, na, nl) for i, (na, nl) in enumerate(sizes)]
    data = synth.batch_from_scenes(scenes)
    oc, orr, oa = ScenePredOracle(ckpt_sd)(data)
    net = make_net(ckpt_sd, dev, "f16tc")
    net.set_option("tc_min_tokens", 0)
    cls, reg, aux = net(to_dev(data, dev))
    torch.cuda.synchronize()
    net.sync_check()
    worst = 0.0
    for b in range(len(sizes)):
        worst = max(worst, rel_err(reg[b], orr[b]), rel_err(aux[b][0], oa[b][0]))
    print("sizes %s fused kernel only: worst rel err %.3e" % ([na + nl + 1 for na, nl in sizes], worst))
    assert worst < 2e-3
