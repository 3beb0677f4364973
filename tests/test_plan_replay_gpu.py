"""GPU: BASELINE.json configs[4] on the B200.  Plan calls recorded while the UNMODIFIED reference drove the demo logs
closed loop on the CPU (oracle/record_plan_calls.py -> tests/golden/plan_calls_demo_*.pt.xz) are replayed on the product's
planner stack (mind_b200/integration/replay.py: scenario tree on the CUDA predictor, cost fields on the GPU, native tree
iLQR, the reference's selection rule).

What is held: every call whose keep / merge decisions are all clear of their threshold (margins recorded from the CPU oracle
tree) reproduces the reference's scenario trees node for node; a call with a decision inside the noise band reproduces
every node created above that depth (the reference's own fp32 evaluation of global-frame coordinates makes such a
decision fall either way -- even the CPU restatement differs from the reference on 10 of the 84 recorded calls).  On the
calls whose trees agree, the control comes out of an iterative optimiser that stops at a relative cost change of 1e-6 and
picks the cheapest of several trajectory trees: predictor inputs that differ in the fifth digit move most controls by
< 1e-3 and a few by more (a different line-search step or a near-tie between two trees), so the test holds the bulk
(>= 80 % of those calls inside TOL_CTRL, median well inside) and prints the tail."""
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
    diffs, same, clear, other_tree = [], 0, 0, 0
    for r in out:
        d0 = fragile_depth(r)
        clear += d0 is None
        if not r["same_trees"]:
            assert d0 is not None, "plan %d: scenario trees differ although every decision is clear of its threshold" % r["plan_index"]
            assert keys_above(r["got_keys"], d0) == keys_above(r["ref_keys"], d0), "plan %d: nodes above depth %d differ" % (r["plan_index"], d0)
            continue
        same += 1
        if r["best_idx"] not in r["ref_best"]:
            other_tree += 1
            continue
        diffs.append(np.abs(r["ctrl"] - r["ref_ctrl"]))
    diffs = np.array(diffs)
    inside = (diffs < TOL_CTRL).all(axis=1)
    ms = 1e3 * np.median([r["seconds"]["total"] for r in out])
    print("%s %s: %d plan calls (%d with every decision clear of its threshold), %d with the reference's trees node for node, %d of "
          "those chose another trajectory tree; |ctrl - reference| median (%.1e m/s^2, %.1e rad/s), max (%.1e, %.1e), %d of %d inside "
          "tolerance; %.1f ms per call" % (rec["demo"], prec, len(out), clear, same, other_tree, *np.median(diffs, axis=0), *diffs.max(axis=0),
                                          inside.sum(), len(diffs), ms))
    assert same >= 0.7 * len(out)
    assert other_tree <= 0.2 * same
    assert inside.sum() >= 0.8 * len(diffs)
    assert (np.median(diffs, axis=0) < 0.2 * TOL_CTRL).all()
