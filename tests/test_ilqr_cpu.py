"""CPU: the native tree iLQR (mind_ilqr_tree_solve, csrc/ilqr_tree.cpp; host code) vs the solutions of the UNMODIFIED
reference optimiser on the demo_2 scenario trees (tests/golden/ilqr_demo_2.npz, oracle/make_golden_ilqr.py):
warm-start solve from zero controls, then the full solve started from the warm-start controls.  Cost fields come from the
cost-field oracle (numpy), so no GPU is involved.  Tolerance 1e-6 on states / controls (fp64; BLAS vs plain-loop
summation orders differ in the last bits and the solver runs tens of iterations)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_cost_field_cpu import CFG, demo2_tree_objects, demo2_trees

W_DES = np.diag([0, 0, 0.1, 0, 1.0, 10.0])                          # planning/demo_*.py:21-24 / 49-52
W_CON = np.diag([0, 0, 50.0, 0, 50.0, 500.0])
UPPER = np.array([100000.0, 100000.0, 8.0, 10.0, 4.0, 0.2])
LOWER = np.array([-100000.0, -100000.0, 0.0, -10.0, -6.0, -0.2])


def opt_cfg(g):
    return dict(CFG, smooth_grid_size=tuple(int(v) for v in g["grid"]), smooth_grid_res=float(g["res"]), w_des_state=W_DES,
                w_state_con=W_CON, state_upper_bound=UPPER, state_lower_bound=LOWER, w_ctrl=5.0 * np.eye(2))


def golden():
    return dict(np.load(os.path.join(GOLDEN, "ilqr_demo_2.npz")))


def fields_from_oracle(nodes, root, tree, x0, lane, cfg, warm):
    from oracle import cost_field_oracle as O
    off, xx, yy, fields, links = O.cost_fields(nodes, root, x0, lane, cfg, warm=warm)
    probs = [p for _, _, p, _, _ in O.walk(nodes, root)]
    return dict(offset=off, xx=xx, yy=yy, fields=fields, links=links, probs=probs)


@pytest.mark.parametrize("ti", [0, 1, 2])
def test_native_ilqr_vs_reference(ti):
    from mind_b200.traj_opt import solve_tree
    g = golden()
    cfg = opt_cfg(g)
    x0 = np.concatenate([g["state"], g["ctrl"]])
    (root, nodes), tree = demo2_trees()[ti], demo2_tree_objects()[ti]
    fw = fields_from_oracle(nodes, root, tree, x0, g["lane"], cfg, True)
    xs_w, us_w, info_w = solve_tree(tree, x0, g["lane"], float(g["target_vel"]), cfg, float(g["dt"]), warm=True, fields=fw)
    assert np.abs(xs_w - g["t%d/warm/xs" % ti]).max() < 1e-6 and np.abs(us_w - g["t%d/warm/us" % ti]).max() < 1e-6
    ff = fields_from_oracle(nodes, root, tree, x0, g["lane"], cfg, False)
    xs, us, info = solve_tree(tree, x0, g["lane"], float(g["target_vel"]), cfg, float(g["dt"]), us_init=g["t%d/warm/us" % ti], warm=False, fields=ff)
    print("tree %d: warm %d iterations, full %d iterations, max |dx| %.2e |du| %.2e" %
          (ti, info_w["iterations"], info["iterations"], np.abs(xs - g["t%d/full/xs" % ti]).max(), np.abs(us - g["t%d/full/us" % ti]).max()))
    assert np.abs(xs - g["t%d/full/xs" % ti]).max() < 1e-6 and np.abs(us - g["t%d/full/us" % ti]).max() < 1e-6


def test_ilqr_argument_checks():
    from mind_b200.traj_opt import ilqr_solve
    n = 3
    ok = dict(parent=[-1, 0, 1], x0=np.zeros(6), dt=0.2, offset=np.zeros(2), xs_grid=np.arange(4.0), ys_grid=np.arange(4.0), res=1.0,
              fields=np.zeros((n, 4, 4)), w_state=np.zeros((n, 6, 6)), des_state=np.zeros((n, 6)), w_con=np.zeros((n, 6, 6)),
              lower=-np.ones(6) * 1e5, upper=np.ones(6) * 1e5, w_ctrl=np.tile(np.eye(2), (n, 1, 1)), us_init=np.zeros((n, 2)))
    xs, us, it, cost = ilqr_solve(**ok)
    assert xs.shape == (3, 6) and np.abs(us).max() == 0 and cost == 0.0        # zero cost everywhere: zero controls are optimal
    with pytest.raises(RuntimeError):
        ilqr_solve(**dict(ok, parent=[-1, 2, 0]))                              # a parent must precede its child
    with pytest.raises(RuntimeError):
        ilqr_solve(**dict(ok, parent=[0, 0, 1]))                               # node 0 hangs off the root state


def test_optimizer_class_surface(monkeypatch):
    """TrajectoryTreeOptimizerB200 keeps the reference class's call sequence (planner.py:171-175) and returns the same
    tree of [state, control] nodes; the field source is redirected to the numpy oracle here (the product uses the GPU)."""
    import types
    from mind_b200 import traj_opt as TO
    from oracle import cost_field_oracle as O
    g = golden()
    (root, nodes), tree = demo2_trees()[1], demo2_tree_objects()[1]

    def oracle_fields(scen_tree, x0, lane, cfg, device, warm=False):
        off, xx, yy, fields, links = O.cost_fields(nodes, root, x0, lane, cfg, warm=warm)
        return dict(offset=off, xx=xx, yy=yy, fields=fields, links=links, probs=[p for _, _, p, _, _ in O.walk(nodes, root)])
    monkeypatch.setattr(TO.CF, "cost_fields", oracle_fields)
    base = opt_cfg(g)
    config = types.SimpleNamespace(dt=float(g["dt"]), state_size=6, action_size=2, w_opt_cfg=dict(base), opt_cfg=dict(base))
    opt = TO.TrajectoryTreeOptimizerB200(config, device="cpu")
    opt.init_warm_start_cost_tree(tree, g["state"], g["ctrl"], g["lane"], float(g["target_vel"]))
    xs, us = opt.warm_start_solve()
    assert np.abs(us - g["t1/warm/us"]).max() < 1e-6
    opt.init_cost_tree(tree, g["state"], g["ctrl"], g["lane"], float(g["target_vel"]))
    tt = opt.solve(us)
    n = len(g["t1/full/xs"])
    assert sorted(tt.nodes, key=lambda k: (k != -1, k)) == [-1] + list(range(n))
    assert tt.nodes[-1].parent_key is None and np.array_equal(tt.nodes[-1].data[0], np.concatenate([g["state"], g["ctrl"]]))
    assert tt.nodes[0].parent_key == -1
    for k in range(n):
        assert np.abs(tt.nodes[k].data[0] - g["t1/full/xs"][k]).max() < 1e-6 and np.abs(tt.nodes[k].data[1] - g["t1/full/us"][k]).max() < 1e-6


def _field_eval(field, xs, ys, res, off, x, y):
    import ctypes as C
    from mind_b200 import lib
    L = lib.load()
    f = np.ascontiguousarray(field[None], dtype=np.float64)
    xs, ys, off = (np.ascontiguousarray(a, dtype=np.float64) for a in (xs, ys, off))
    p = lib.MindIlqrTree()
    p.n_nodes, p.gx, p.gy, p.res = 1, len(xs), len(ys), float(res)
    p.fields, p.xs_grid, p.ys_grid, p.field_offset = f.ctypes.data, xs.ctypes.data, ys.ctypes.data, off.ctypes.data
    out = (C.c_double * 6)()
    assert L.mind_debug_field_eval(C.byref(p), 0, float(x), float(y), out) == 0
    return np.array(out[:])


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(), reason="needs the reference tree (build container only)")
def test_field_patch_vs_reference_potential_field():
    """interior, edge and corner cells, positions outside the grid: value / gradient / Hessian of the native patch code vs
    the reference's PotentialField (potential.py:62-264), including its zero-padded border neighbourhoods"""
    import sys
    from oracle import ref_loader
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    from planners.ilqr.potential import PotentialField
    rng = np.random.default_rng(3)
    gx, gy, res = 9, 7, 0.8
    off = np.array([100.0, -50.0])
    xs, ys = np.linspace(0, (gx - 1) * res, gx) + off[0], np.linspace(0, (gy - 1) * res, gy) + off[1]
    xx, yy = np.meshgrid(xs, ys)
    field = rng.uniform(0, 10, size=(gy, gx))
    pf = PotentialField(off, res, xx, yy, field)
    pts = [(xs[i] + dx, ys[j] + dy) for i in (0, 1, 4, gx - 2, gx - 1) for j in (0, 1, 3, gy - 2, gy - 1)
           for dx, dy in ((0.0, 0.0), (0.3, -0.25), (-0.39, 0.39))]
    pts += [(xs[0] - 5.0, ys[2]), (xs[-1] + 7.0, ys[-1] + 3.0), (xs[3] + 0.4, ys[3] - 0.4), (xs[2] + 0.4000001, ys[2])]
    worst = 0.0
    for x, y in pts:
        st = np.array([x, y, 1.0, 0.2, 0.0, 0.0])
        want = np.concatenate([[pf.get_potential(st)], pf.get_gradient(st)[:2],
                               [pf.get_hessian(st)[0, 0], pf.get_hessian(st)[0, 1], pf.get_hessian(st)[1, 1]]])
        got = _field_eval(field, xs, ys, res, off, x, y)
        worst = max(worst, np.abs(got - want).max() / max(1.0, np.abs(want).max()))
    assert worst < 1e-13, worst


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_native_ilqr_vs_live_reference_on_random_trees(seed):
    """random tree shapes, random rough cost fields, tight state bounds (constraint potentials active), states that leave
    the grid: the reference's iLQR.fit on a TreeCost built from its own potentials vs mind_ilqr_tree_solve"""
    import sys
    from oracle import ref_loader
    from mind_b200 import compat
    compat.install()
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    from planners.basic.tree import Tree, Node
    from planners.ilqr.cost import TreeCost
    from planners.ilqr.potential import ControlPotential, PotentialField, StateConstraint, StatePotential
    from planners.mind.trajectory_tree import TrajectoryTreeOptimizer
    from planners.mind.configs.planning.demo_1 import TrajTreeCfg
    from mind_b200.traj_opt import ilqr_solve
    rng = np.random.default_rng(seed)
    n = int(rng.integers(6, 30))
    parent = [-1] + [int(rng.integers(max(0, i - 4), i)) for i in range(1, n)]
    gx, gy, res = int(rng.integers(12, 40)), int(rng.integers(12, 40)), float(rng.choice([0.4, 1.0, 2.5]))
    x0 = np.array([50.0, -20.0, float(rng.uniform(0, 9)), float(rng.uniform(-3, 3)), float(rng.uniform(-1, 1)), float(rng.uniform(-0.1, 0.1))])
    off = x0[:2] - 0.5 * np.array([(gx - 1) * res, (gy - 1) * res])
    xs_g, ys_g = np.linspace(0, (gx - 1) * res, gx) + off[0], np.linspace(0, (gy - 1) * res, gy) + off[1]
    xx, yy = np.meshgrid(xs_g, ys_g)
    base = ((xx - x0[0] - 3.0) ** 2 + (yy - x0[1] + 2.0) ** 2) * float(rng.uniform(0.2, 2.0))
    fields = base[None] * rng.uniform(0.3, 1.0, size=(n, 1, 1)) + rng.uniform(0, 3.0, size=(n, gy, gx))
    probs = rng.uniform(0.2, 1.0, size=n)
    w_des = np.diag([0, 0, 0.1, 0, 1.0, 10.0]); w_con = np.diag([0, 0, 50.0, 0, 50.0, 500.0]); w_u = 5.0 * np.eye(2)
    upper = np.array([1e5, 1e5, float(rng.uniform(3, 8)), 10.0, 1.0, 0.05]); lower = np.array([-1e5, -1e5, 0.5, -10.0, -1.0, -0.05])
    des = np.array([0, 0, float(rng.uniform(2, 9)), 0.0, 0.0, 0.0])
    tree = Tree()
    tree.add_node(Node(-1, None, x0))
    for i in range(n):
        pots = [PotentialField(off, res, xx, yy, fields[i]), StatePotential(w_des * probs[i], des), StateConstraint(w_con * probs[i], lower, upper)]
        tree.add_node(Node(i, parent[i], [pots, [ControlPotential(w_u * probs[i])]]))
    opt = TrajectoryTreeOptimizer(TrajTreeCfg())
    us0 = rng.normal(size=(n, 2)) * np.array([0.5, 0.02])
    xs_r, us_r = opt.ilqr.fit(us0, TreeCost(tree, 6, 2))
    xs, us, it, cost = ilqr_solve(parent, x0, 0.2, off, xs_g, ys_g, res, fields, probs[:, None, None] * w_des[None], np.tile(des, (n, 1)),
                                  probs[:, None, None] * w_con[None], lower, upper, probs[:, None, None] * w_u[None], us0)
    err = max(np.abs(xs - xs_r).max(), np.abs(us - us_r).max())
    print("seed %d: %d nodes, grid %dx%d @ %.1f m, %d iterations, max diff %.2e" % (seed, n, gx, gy, res, it, err))
    assert err < 1e-6
