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


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_cost_fields_vs_live_reference_on_random_trees(seed):
    """random scenario trees (1..6 actors incl. ego-only, odd and even durations, branching), random lanes and grids:
    the unmodified TrajectoryTreeOptimizer.init_warm_start_cost_tree / init_cost_tree vs the oracle and vs the host
    tables of mind_b200.cost_field"""
    import sys
    from oracle import ref_loader, cost_field_oracle as O
    from mind_b200 import compat
    from mind_b200.cost_field import node_tables
    compat.install()
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    from planners.basic.tree import Tree, Node
    from planners.mind.trajectory_tree import TrajectoryTreeOptimizer
    from planners.mind.configs.planning.demo_3 import TrajTreeCfg
    rng = np.random.default_rng(100 + seed)
    na = int(rng.integers(1, 7))
    keys, parents = ["r"], {"r": None}
    for i in range(int(rng.integers(1, 6))):
        k = "n%d" % i
        parents[k] = keys[int(rng.integers(0, len(keys)))]
        keys.append(k)
    flat = {}
    for k in keys:
        dur = int(rng.integers(1, 12))
        trajs = (rng.normal(size=(na, dur, 2)) * 8 + np.array([100.0, 40.0])).astype(np.float32)
        covs = rng.uniform(0.0, 2.0, size=(na, dur, 1)).astype(np.float32)
        flat[k] = (parents[k], float(rng.uniform(0.05, 1.0)), trajs, covs)
    tree = Tree()
    for k in keys:
        tree.add_node(Node(k, parents[k], [flat[k][1], flat[k][2], flat[k][3], None]))
    nodes = {k: (parents[k], flat[k][1], flat[k][2], flat[k][3], list(tree.get_node(k).children_keys)) for k in keys}
    cfgobj = TrajTreeCfg()
    grid, res = (int(rng.integers(8, 24)), int(rng.integers(8, 24))), float(rng.choice([0.4, 1.0, 3.0]))
    for c in (cfgobj.w_opt_cfg, cfgobj.opt_cfg):
        c["smooth_grid_size"], c["smooth_grid_res"] = grid, res
    lane = np.cumsum(rng.normal(size=(int(rng.integers(2, 12)), 2)) * 6, axis=0) + np.array([95.0, 35.0])
    state, ctrl = np.array([100.0, 40.0, 5.0, 0.3]), np.array([0.1, 0.0])
    opt = TrajectoryTreeOptimizer(cfgobj)
    x0 = np.concatenate([state, ctrl])
    for warm, init, cfg in ((True, opt.init_warm_start_cost_tree, cfgobj.w_opt_cfg), (False, opt.init_cost_tree, cfgobj.opt_cfg)):
        init(tree, state, ctrl, lane, 7.0)
        rn = opt.cost_tree.tree.nodes
        rkeys = [k for k in rn if k != -1]
        want = np.stack([rn[k].data[0][0].cost_field for k in rkeys])
        off, xx, yy, got, links = O.cost_fields(nodes, "r", x0, lane, cfg, warm=warm)
        assert links == [(k, rn[k].parent_key) for k in rkeys]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
        coef, mean, rad, plinks, _ = node_tables(tree, cfg, warm)
        o_coef, o_mean, o_rad, _ = O.node_inputs(nodes, "r", cfg, warm)
        assert plinks == links and np.array_equal(coef, o_coef)
        if not warm:
            assert np.array_equal(mean, o_mean) and np.array_equal(rad, o_rad)
