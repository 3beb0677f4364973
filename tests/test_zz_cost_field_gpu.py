"""GPU: mind_cost_fields (csrc/cost_field.cu) vs the oracle and vs the golden fields of the unmodified reference optimiser
(tests/golden/cost_fields_demo_2.npz), fp64, 1e-12 relative; at the reference's full 256 x 256 grid through
size-independent properties.  First hardware run (profiles/r01_v8_cost_field_gpu_tests.log): warm-start fields bit-exact,
full fields within 1.8e-16 relative of the reference's."""
import numpy as np
import pytest
import torch

from test_cost_field_cpu import cfg_of, demo2_tree_objects, demo2_trees, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("warm", [True, False])
def test_cost_fields_vs_reference_golden(warm):
    from mind_b200.cost_field import cost_fields
    g = golden()
    x0 = np.concatenate([g["state"], g["ctrl"]])
    for ti, tree in enumerate(demo2_tree_objects()):
        tag = "t%d/%s" % (ti, "warm" if warm else "full")
        out = cost_fields(tree, x0, g["lane"], cfg_of(g), torch.device("cuda", 0), warm=warm)
        want = g[tag + "/fields"]
        assert np.array_equal(np.array(out["links"]), g[tag + "/links"]) and np.array_equal(out["offset"], g[tag + "/offset"])
        assert out["fields"].shape == want.shape
        err = np.abs(out["fields"] - want).max() / np.abs(want).max()
        print("tree %d %s: %d fields, max rel err %.2e" % (ti, tag, len(want), err))
        assert err <= 1e-12


def test_cost_fields_full_grid_properties():
    """256 x 256 cells, 0.4 m (demo_*.py:45-46,73-74): vs the oracle on one tree, plus properties that hold at any size"""
    from mind_b200.cost_field import cost_fields
    from oracle import cost_field_oracle as O
    g = golden()
    cfg = dict(cfg_of(g), smooth_grid_size=(256, 256), smooth_grid_res=0.4)
    x0 = np.concatenate([g["state"], g["ctrl"]])
    (root, nodes), tree = demo2_trees()[0], demo2_tree_objects()[0]
    full = cost_fields(tree, x0, g["lane"], cfg, torch.device("cuda", 0), warm=False)
    warm = cost_fields(tree, x0, g["lane"], cfg, torch.device("cuda", 0), warm=True)
    assert full["fields"].shape == (len(full["links"]), 256, 256)
    assert (full["fields"] >= warm["fields"] - 1e-9).all()                         # ego / exo terms only add cost
    assert np.allclose(warm["fields"], np.array([1.0 * p for p in warm["probs"]])[:, None, None] * warm["quad"][None], rtol=1e-15)
    assert np.isfinite(full["fields"]).all() and warm["quad"].min() >= 0.0
    _, _, _, want, links = O.cost_fields(nodes, root, x0, g["lane"], cfg, warm=False)
    assert links == full["links"]
    assert np.abs(full["fields"] - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("ti", [0, 1, 2])
def test_optimizer_on_gpu_fields_vs_reference_solution(ti):
    """the whole accelerated optimiser step: cost fields on the GPU (mind_cost_fields) + native tree iLQR, against the
    solutions of the unmodified reference optimiser (tests/golden/ilqr_demo_2.npz); warm start, then the full solve"""
    from mind_b200.traj_opt import solve_tree
    from test_ilqr_cpu import golden as ilqr_golden, opt_cfg
    g = ilqr_golden()
    cfg = opt_cfg(g)
    x0 = np.concatenate([g["state"], g["ctrl"]])
    tree = demo2_tree_objects()[ti]
    dev = torch.device("cuda", 0)
    # tolerance: the solver stops at a relative cost change of 1e-6, so its solution is defined to ~1e-3 in the states; with
    # fields that differ from the reference's in the last bit (<= 1.8e-16) the iterates normally coincide to ~1e-11 (the CPU
    # test holds 1e-6 with oracle fields), but one flipped accept decision would move them by ~1e-3
    tol = 1e-2
    xs_w, us_w, _ = solve_tree(tree, x0, g["lane"], float(g["target_vel"]), cfg, float(g["dt"]), warm=True, device=dev)
    assert np.abs(xs_w - g["t%d/warm/xs" % ti]).max() < tol and np.abs(us_w - g["t%d/warm/us" % ti]).max() < tol
    xs, us, info = solve_tree(tree, x0, g["lane"], float(g["target_vel"]), cfg, float(g["dt"]), us_init=g["t%d/warm/us" % ti], warm=False, device=dev)
    print("tree %d: %d iterations, max |dx| %.2e (warm %.2e)" % (ti, info["iterations"], np.abs(xs - g["t%d/full/xs" % ti]).max(),
                                                                  np.abs(xs_w - g["t%d/warm/xs" % ti]).max()))
    assert np.abs(xs - g["t%d/full/xs" % ti]).max() < tol and np.abs(us - g["t%d/full/us" % ti]).max() < tol
