"""Closed-loop run of the UNMODIFIED reference planner stack on a demo log (BASELINE.json configs[4]), without
rendering:   python -m oracle.run_closed_loop demo_2 [--horizon 6.0]

Build container only (needs /root/reference; the reference sources cannot travel to the GPU box).  av2 / shapely /
Theano come from mind_b200.compat, so this doubles as the end-to-end check of those stand-ins: the reference's loader,
SemanticMap, agents, MINDPlanner (scenario tree on its own CPU ScenePredNet + tree iLQR on the theano_lite bicycle model)
run exactly as simulator.py:52-107 drives them.  Prints per-plan wall time split and a few sanity figures of the drive.
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("demo")
    ap.add_argument("--horizon", type=float, default=6.0, help="simulated seconds (reference: 10.0)")
    ap.add_argument("--native-ilqr", action="store_true",
                    help="swap the reference's numpy tree iLQR for mind_ilqr_tree_solve (cost fields from the numpy oracle: no GPU here)")
    args = ap.parse_args()
    from oracle import ref_loader
    from mind_b200 import compat
    print("stand-ins:", compat.install())
    sys.path.insert(0, ref_loader.REF_ROOT)
    from pathlib import Path
    import torch
    from common.semantic_map import SemanticMap
    from loader import ArgoAgentLoader
    from agent import CustomizedAgent, NonReactiveAgent
    cfg = json.load(open(os.path.join(ref_loader.REF_ROOT, "configs", args.demo + ".json")))
    seq = cfg["seq_id"]
    seq_path = os.path.join(ref_loader.REF_ROOT, "data", seq)
    smp = SemanticMap()
    smp.load_from_argo2(Path(os.path.join(seq_path, "log_map_archive_%s.json" % seq)))
    os.chdir(ref_loader.REF_ROOT)
    agents = ArgoAgentLoader(Path(os.path.join(seq_path, "scenario_%s.parquet" % seq))).load_agents(smp, cfg["cl_agents"])
    ego = [a for a in agents if isinstance(a, CustomizedAgent)][0]
    # wall-time split of plan(): wrap the two halves of MINDPlanner.plan (planner.py:104-125) without changing them
    pl = ego.planner
    t_tree, t_ilqr = [], []
    branch_aime, get_traj_tree = pl.scen_tree_gen.branch_aime, pl.get_traj_tree

    def timed_branch(lcl_smp, obs):
        t0 = time.perf_counter()
        with torch.no_grad():
            out = branch_aime(lcl_smp, obs)
        t_tree.append(time.perf_counter() - t0)
        return out

    def timed_traj(scen_tree, lcl_smp):
        t0 = time.perf_counter()
        out = get_traj_tree(scen_tree, lcl_smp)
        t_ilqr.append(time.perf_counter() - t0)
        return out
    pl.scen_tree_gen.branch_aime, pl.get_traj_tree = timed_branch, timed_traj
    if args.native_ilqr:
        # the product's optimiser class with the field source redirected to the CPU oracle (test infrastructure: the
        # product computes the fields with mind_cost_fields on the GPU)
        from mind_b200 import traj_opt as TO
        from oracle import cost_field_oracle as O

        def oracle_fields(scen_tree, x0, lane, cfg, device, warm=False):
            nodes = {k: (n.parent_key, n.data[0], n.data[1], n.data[2], list(n.children_keys)) for k, n in scen_tree.nodes.items()}
            root = scen_tree.get_root().key
            off, xx, yy, fields, links = O.cost_fields(nodes, root, x0, lane, cfg, warm=warm)
            return dict(offset=off, xx=xx, yy=yy, fields=fields, links=links, probs=[p for _, _, p, _, _ in O.walk(nodes, root)])
        TO.CF.cost_fields = oracle_fields
        pl.traj_tree_opt = TO.TrajectoryTreeOptimizerB200(pl.traj_tree_opt.config, device="cpu")
    # finer split of the iLQR half (trajectory_tree.py:20-124 cost trees vs :125-147 solves)
    opt, t_parts = pl.traj_tree_opt, {}
    for meth in ("init_warm_start_cost_tree", "init_cost_tree", "warm_start_solve", "solve"):
        def wrap(fn, key):
            def run(*a, **k):
                t0 = time.perf_counter()
                out = fn(*a, **k)
                t_parts[key] = t_parts.get(key, 0.0) + time.perf_counter() - t0
                return out
            return run
        setattr(opt, meth, wrap(getattr(opt, meth), meth))
    sim_time, step, plans, n_trees = 0.0, 0.02, 0, []
    min_gap, lane_dev = 1e9, []
    t_start = time.perf_counter()
    while sim_time < args.horizon:
        obs = [a.observe() for a in agents if (isinstance(a, NonReactiveAgent) and a.is_valid()) or isinstance(a, CustomizedAgent)]
        for a in agents:
            if isinstance(a, CustomizedAgent):
                a.check_enable(sim_time)
                rec_tri, pl_tri = a.check_trigger(sim_time)
                if rec_tri:
                    a.step()
                if pl_tri:
                    a.update_observation(obs)
                    if a.is_enable:
                        ok, res = a.plan()
                        assert ok, "plan failed at t=%.2f" % sim_time
                        plans += 1
                        n_trees.append(len(t_ilqr))
            else:
                a.step()
            a.update_state(step)
        if ego.is_enable:
            others = [o.state[:2] for o in obs if o.id != ego.id]
            if others:
                min_gap = min(min_gap, float(np.min(np.linalg.norm(np.array(others) - ego.state[:2], axis=1))))
            lane_dev.append(pl.get_dist_to_target_lane(ego.lcl_smp, ego.state))
        sim_time += step
    wall = time.perf_counter() - t_start
    print("%s: %.1f s simulated in %.1f s wall, %d plan calls (%.0f ms each: scenario tree %.0f ms, tree iLQR %.0f ms over %.1f trees)" %
          (args.demo, sim_time, wall, plans, 1e3 * (sum(t_tree) + sum(t_ilqr)) / max(plans, 1), 1e3 * np.mean(t_tree),
           1e3 * sum(t_ilqr) / max(plans, 1), len(t_ilqr) / max(plans, 1)))
    print("iLQR half per plan call [ms]:", {k: round(1e3 * v / max(plans, 1), 1) for k, v in t_parts.items()})
    print("ego final state (x, y, v, heading):", np.round(ego.state, 3), "| target velocity", ego.lcl_smp.target_velocity)
    print("closest other agent while enabled: %.2f m; distance to the target lane: mean %.2f m, max %.2f m" %
          (min_gap, float(np.mean(lane_dev)), float(np.max(lane_dev))))


if __name__ == "__main__":
    main()
