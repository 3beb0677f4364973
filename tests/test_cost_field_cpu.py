"""CPU: the cost-field oracle (oracle/cost_field_oracle.py) vs the fields the UNMODIFIED reference optimiser built on the
demo_2 scenario trees (tests/golden/cost_fields_demo_2.npz, oracle/make_golden_cost_fields.py): creation order and parent
links exact, values to 1e-12 relative (fp64; numpy's dgemv / norm vs element-wise forms differ in the last bits)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

CFG = dict(w_tgt=1.0, w_ego=1.0, w_ego_cov_offset=1.0, w_exo=10.0, w_exo_cov_offset=2.5, w_exo_cost_offset=10.0)   # demo_*.py:71-81


def demo2_trees():
    """[(root key, {key: (parent, prob, trajs, covs, children)})] of the golden demo_2 scenario trees"""
    flat = torch.load(os.path.join(GOLDEN, "real_demo_2.pt"), weights_only=False)["tree"]
    kids = {}
    for k, v in flat.items():
        kids.setdefault(v[0], []).append(k)
    nodes = {k: (v[0], v[1], v[2], v[3], sorted(kids.get(k, []))) for k, v in flat.items()}
    return [(rk, nodes) for rk in sorted(k for k, v in flat.items() if v[0] is None)]


def golden():
    return dict(np.load(os.path.join(GOLDEN, "cost_fields_demo_2.npz")))


def cfg_of(g):
    return dict(CFG, smooth_grid_size=tuple(int(v) for v in g["grid"]), smooth_grid_res=float(g["res"]))


@pytest.mark.parametrize("warm", [True, False])
def test_oracle_cost_fields_vs_reference(warm):
    from oracle.cost_field_oracle import cost_fields
    g = golden()
    x0 = np.concatenate([g["state"], g["ctrl"]])
    for ti, (root, nodes) in enumerate(demo2_trees()):
        tag = "t%d/%s" % (ti, "warm" if warm else "full")
        off, xx, yy, fields, links = cost_fields(nodes, root, x0, g["lane"], cfg_of(g), warm=warm)
        want = g[tag + "/fields"]
        assert np.array_equal(np.array(links), g[tag + "/links"])
        assert np.abs(off - g[tag + "/offset"]).max() == 0 and fields.shape == want.shape and xx.shape == (20, 28)
        assert np.abs(fields - want).max() <= 1e-12 * np.abs(want).max(), np.abs(fields - want).max()
        if not warm:
            assert (fields >= g["t%d/warm/fields" % ti] - 1e-9).all()      # obstacle / ego terms only add cost


def test_walk_order_is_lifo_depth_first():
    from oracle.cost_field_oracle import walk
    z = np.zeros((1, 3, 2))
    nodes = {"a": (None, 1.0, z, z, ["b", "c"]), "b": ("a", 0.5, z, z, []), "c": ("a", 0.5, z, z, ["d"]), "d": ("c", 0.5, z[:, :1], z, [])}
    got = [(k, i, idx, last) for k, i, _, idx, last in walk(nodes, "a")]
    assert got == [("a", 0, 0, -1), ("a", 2, 1, 0), ("c", 0, 2, 1), ("c", 2, 3, 2), ("d", 0, 4, 3), ("b", 0, 5, 1), ("b", 2, 6, 5)]


class _Node:
    def __init__(self, key, parent_key, data):
        self.key, self.parent_key, self.data, self.children_keys = key, parent_key, data, []


class _Tree:            # the slice of planners/basic/tree.py the cost-field walk uses
    def __init__(self, root, nodes):
        self.nodes = {}
        order, seen = [root], 0
        while seen < len(order):
            k = order[seen]; seen += 1
            parent, prob, trajs, covs, children = nodes[k]
            self.nodes[k] = _Node(k, parent, [prob, trajs, covs, None])
            self.nodes[k].children_keys = list(children)
            order += children
        self.root = root

    def get_root(self):
        return self.nodes[self.root]

    def get_node(self, k):
        return self.nodes[k]


def demo2_tree_objects():
    return [_Tree(root, nodes) for root, nodes in demo2_trees()]


@pytest.mark.parametrize("warm", [True, False])
def test_product_host_tables_match_the_oracle(warm):
    """host half of mind_b200.cost_field (tree walk, coefficient / centre / radius tables, grid frame): no GPU needed"""
    from mind_b200.cost_field import grid_frame, node_tables
    from oracle import cost_field_oracle as O
    g = golden()
    cfg = cfg_of(g)
    x0 = np.concatenate([g["state"], g["ctrl"]])
    off, xs, ys = grid_frame(x0, cfg["smooth_grid_size"], cfg["smooth_grid_res"])
    o_off, xx, yy = O.grid_frame(x0, cfg["smooth_grid_size"], cfg["smooth_grid_res"])
    assert np.array_equal(off, o_off) and np.array_equal(xs, xx[0]) and np.array_equal(ys, yy[:, 0])
    for (root, nodes), tree in zip(demo2_trees(), demo2_tree_objects()):
        coef, mean, rad, links, _ = node_tables(tree, cfg, warm)
        o_coef, o_mean, o_rad, o_links = O.node_inputs(nodes, root, cfg, warm)
        assert np.array_equal(coef, o_coef) and links == o_links
        if not warm:
            assert np.array_equal(mean, o_mean) and np.array_equal(rad, o_rad)
        else:
            assert mean is None and rad is None
