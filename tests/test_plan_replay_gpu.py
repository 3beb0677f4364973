"""GPU: BASELINE.json configs[4] on the B200.  Plan calls recorded while the UNMODIFIED reference drove the demo logs
closed loop on the CPU (oracle/record_plan_calls.py -> tests/golden/plan_calls_demo_*.pt.xz) are replayed on the product's
planner stack (mind_b200/integration/replay.py: scenario tree on the CUDA predictor, cost fields on the GPU, native tree
iLQR, the reference's selection rule).

What is held: every call whose keep / merge decisions are all clear of their threshold (margins recorded from the CPU oracle
tree) reproduces the reference's scenario trees node for node; a call with a decision inside the noise band reproduces
every node created above that depth (the reference's own fp32 evaluation of global-frame coordinates makes such a
decision fall either way -- even the CPU restatement differs from the reference on some of them); whenever the trees are
the same, the chosen tree and the control are the reference's."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

FILES = sorted(glob.glob(os.path.join(GOLDEN, "plan_calls_demo_*.pt.xz")))
# acceleration [m/s^2], steering rate [rad/s]: the reference's own solver stops at a relative cost change of 1e-6, so two
# implementations of it agree to about 1e-3 of the control range (tests/test_ilqr_cpu.py pins single solves to 5e-11)
TOL_CTRL = np.array([5e-3, 1e-3])


@pytest.mark.parametrize("prec", ["f16tc", "fp32"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p).split(".")[0] for p in FILES])
def test_replayed_plan_calls_give_the_reference_controls(ckpt_sd, path, prec):
    from mind_b200.integration.replay import fragile_depth, keys_above, replay_file
    from mind_b200.predictor import ScenePredNetB200
    dev = torch.device("cuda", 0)
    net = ScenePredNetB200(None, dev)
    net.load_state_dict(ckpt_sd)
    net.set_precision(prec)
    rec, out = replay_file(path, dev, net)
    assert len(out) >= 1
    worst, same, clear = np.zeros(2), 0, 0
    for r in out:
        d0 = fragile_depth(r)
        clear += d0 is None
        if not r["same_trees"]:
            assert d0 is not None, "plan %d: scenario trees differ although every decision is clear of its threshold" % r["plan_index"]
            assert keys_above(r["got_keys"], d0) == keys_above(r["ref_keys"], d0), "plan %d: nodes above depth %d differ" % (r["plan_index"], d0)
            continue
        same += 1
        assert r["best_idx"] in r["ref_best"], "plan %d: chose tree %d, reference chose %s" % (r["plan_index"], r["best_idx"], r["ref_best"])
        worst = np.maximum(worst, np.abs(r["ctrl"] - r["ref_ctrl"]))
    ms = 1e3 * np.median([r["seconds"]["total"] for r in out])
    print("%s %s: %d plan calls (%d with every decision clear of its threshold), %d with the reference's trees node for node; on those "
          "max |ctrl - reference| = (%.2e m/s^2, %.2e rad/s); %.1f ms per call" % (rec["demo"], prec, len(out), clear, same, worst[0], worst[1], ms))
    assert same >= 1
    assert (worst < TOL_CTRL).all(), worst
