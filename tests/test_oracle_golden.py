"""CPU: the oracle restatement reproduces the golden vectors dumped from the reference
(tests/golden/*.npz, made by oracle/make_golden.py) and the survey's S1 known answers."""
import numpy as np
import torch

from conftest import load_golden, rel_err
from mind_b200 import synth
from oracle.make_golden import ragged_batch
from oracle.scene_pred_oracle import ScenePredOracle

TOL = 2e-5


def _check(orc, data, gold):
    st = orc.stages(data)
    for k, g in [("actor_feat", "actor_feat"), ("lane_feat", "lane_feat"), ("tgt_feat", "tgt_feat"),
                 ("actors", "actors"), ("lanes", "lanes"), ("cls", "cls_tok")]:
        assert rel_err(st[k], gold[g]) < TOL, k
    cls, reg, aux = orc(data)
    for b in range(len(cls)):
        assert np.abs(cls[b].numpy() - gold["cls_%d" % b]).max() < 1e-6
        assert rel_err(reg[b], gold["reg_%d" % b]) < TOL
        assert rel_err(aux[b][0], gold["vel_%d" % b]) < TOL
        assert rel_err(aux[b][1], gold["covvel_%d" % b]) < TOL
        assert rel_err(aux[b][2], gold["param_%d" % b]) < TOL
        assert (np.argsort(-cls[b].numpy()[0]) == np.argsort(-gold["cls_%d" % b][0])).all()


def test_s1_known_answers(ckpt_sd):
    """values recorded from the reference at survey time."""
    gold = load_golden("s1_ckpt.npz")
    c = gold["cls_0"][0]
    np.testing.assert_allclose(c, [0.071799, 0.266260, 0.071338, 0.055918, 0.467787, 0.066898], atol=2e-6)
    assert list(np.argsort(-c)) == [4, 1, 0, 2, 5, 3]
    assert abs(gold["reg_0"].sum() - 58971.613) < 0.5
    assert abs(np.abs(gold["vel_0"]).sum() - 14568.254) < 0.2
    _check(ScenePredOracle(ckpt_sd), synth.batch_from_scenes([synth.scene_s1(1234)]), gold)


def test_ragged_ckpt(ckpt_sd):
    _check(ScenePredOracle(ckpt_sd), ragged_batch(), load_golden("ragged_ckpt.npz"))


def test_ragged_random_weights(rand_sd):
    _check(ScenePredOracle(rand_sd), ragged_batch(), load_golden("ragged_rand.npz"))


def test_rpe_mirror_matches_oracle():
    from oracle.scene_pred_oracle import get_rpe
    s = synth.scene_s1(5, 4, 7, with_geom=True)
    assert torch.equal(get_rpe(s["ctrs"], s["vecs"]), s["rpe"])
    # diagonal quirk (reference utils.py:205-207,228): zero displacement -> cos/sin 0
    assert float(s["rpe"][2].diagonal().abs().max()) == 0.0
