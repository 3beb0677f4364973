"""Plan-call fixtures of the closed loop (BASELINE.json configs[4]):   python -m oracle.record_plan_calls [demo ...]

TEST INFRASTRUCTURE, build container only (needs /root/reference).  Drives the UNMODIFIED reference stack closed loop
on a demo log exactly like simulator.py:52-107 (same harness as oracle/run_closed_loop.py: its own CPU ScenePredNet,
ScenarioTreeGenerator, TrajectoryTreeOptimizer with the numpy tree iLQR; av2 / shapely / Theano through mind_b200.compat)
and records every `--every`-th call of MINDPlanner.plan (planner.py:104-145):

  inputs    collated scene dict process_data returned (dense RPE dropped: it is get_rpe of the anchors that are kept),
            resampled target lane + info, generator.lane_graph, planner state / ctrl / gt_tgt_lane, target velocity and
            the local target lane (for evaluate_traj_tree, :177-196), the TrajTreeCfg dictionaries
  outputs   the control plan() returned, the index of the chosen tree, node keys / parents of every scenario tree,
            states / controls of every trajectory tree, wall time of the two halves on this container's CPU

The reference sources cannot travel to the GPU box; these records let the product's planner stack (scenario tree on the
CUDA predictor, cost fields on the GPU, native tree iLQR) be replayed there call by call: mind_b200/integration/replay.py,
tests/test_plan_replay_gpu.py, bench.py `closed_loop`.
Writes tests/golden/plan_calls_<demo>.pt.xz (torch.save + lzma).
"""
import argparse
import copy
import io
import json
import lzma
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "tests", "golden")


def cpu_copy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().clone()
    if isinstance(x, dict):
        return {k: cpu_copy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [cpu_copy(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.copy()
    return copy.deepcopy(x)


def record(demo, horizon, every):
    from oracle import ref_loader
    from mind_b200 import compat
    compat.install()
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    from pathlib import Path
    from common.semantic_map import SemanticMap
    from loader import ArgoAgentLoader
    from agent import CustomizedAgent, NonReactiveAgent
    cfg = json.load(open(os.path.join(ref_loader.REF_ROOT, "configs", demo + ".json")))
    seq = cfg["seq_id"]
    seq_path = os.path.join(ref_loader.REF_ROOT, "data", seq)
    smp = SemanticMap()
    smp.load_from_argo2(Path(os.path.join(seq_path, "log_map_archive_%s.json" % seq)))
    cwd = os.getcwd()
    os.chdir(ref_loader.REF_ROOT)
    agents = ArgoAgentLoader(Path(os.path.join(seq_path, "scenario_%s.parquet" % seq))).load_agents(smp, cfg["cl_agents"])
    ego = [a for a in agents if isinstance(a, CustomizedAgent)][0]
    pl = ego.planner
    gen = pl.scen_tree_gen
    cur = {}

    process_data, set_target_lane, branch_aime, get_traj_tree = gen.process_data, gen.set_target_lane, gen.branch_aime, pl.get_traj_tree

    def rec_process(lcl_smp, obs):
        data = process_data(lcl_smp, obs)
        d = cpu_copy(data)
        d.pop("RPE", None)                                   # = get_rpe(cat(TRAJS_CTRS, lane_ctrs), cat(TRAJS_VECS, lane_vecs)), utils.py:193-212
        cur["data"], cur["graph"] = d, cpu_copy(gen.lane_graph)
        return data

    def rec_lane(lane, info):
        cur["lane"], cur["info"] = np.asarray(lane).copy(), [np.asarray(x).copy() for x in info]
        return set_target_lane(lane, info)

    def rec_branch(lcl_smp, obs):
        t0 = time.perf_counter()
        with torch.no_grad():
            trees = branch_aime(lcl_smp, obs)
        cur["t_tree"] = time.perf_counter() - t0
        cur["scen_trees"] = [{k: (n.parent_key, float(n.data[0]), int(n.data[1].shape[1])) for k, n in t.nodes.items()} for t in trees]
        cur["traj_trees"], cur["t_opt"] = [], 0.0
        return trees

    def rec_traj(scen_tree, lcl_smp):
        t0 = time.perf_counter()
        tree, dbg = get_traj_tree(scen_tree, lcl_smp)
        cur["t_opt"] += time.perf_counter() - t0
        cur["traj_trees"].append({k: (n.parent_key, np.asarray(n.data[0], dtype=np.float64).copy(), np.asarray(n.data[1], dtype=np.float64).copy())
                                  for k, n in tree.nodes.items()})
        return tree, dbg
    gen.process_data, gen.set_target_lane, gen.branch_aime, pl.get_traj_tree = rec_process, rec_lane, rec_branch, rec_traj
    plan, records, count = pl.plan, [], [0]

    def rec_plan(lcl_smp):
        cur.clear()
        state, ctrl = np.asarray(pl.state, dtype=np.float64).copy(), np.asarray(pl.ctrl, dtype=np.float64).copy()
        ok, ret_ctrl, res = plan(lcl_smp)
        if ok and count[0] % every == 0:
            best = [i for i, t in enumerate(cur["scen_trees"]) if sorted(t) == sorted(res[0][0].nodes)]
            records.append(dict(sim_time=clock[0], plan_index=count[0], data=cur["data"], graph=cur["graph"],
                                lane=cur["lane"], info=cur["info"], state=state, ctrl=ctrl,
                                gt_tgt_lane=np.asarray(pl.gt_tgt_lane, dtype=np.float64).copy(),
                                target_velocity=float(lcl_smp.target_velocity),
                                lcl_target_lane=np.asarray(lcl_smp.target_lane, dtype=np.float64).copy(),
                                ret_ctrl=np.asarray(ret_ctrl, dtype=np.float64).copy(), best_candidates=best,
                                scen_trees=cur["scen_trees"], traj_trees=cur["traj_trees"],
                                cpu_reference_s=dict(scenario_tree=cur["t_tree"], optimizer=cur["t_opt"])))
            print("%s t=%.2f plan %d: %d trees, ctrl %s, tree %.2f s, optimiser %.2f s" %
                  (demo, clock[0], count[0], len(cur["scen_trees"]), np.round(ret_ctrl, 4), cur["t_tree"], cur["t_opt"]), flush=True)
        count[0] += 1
        return ok, ret_ctrl, res
    pl.plan = rec_plan
    clock = [0.0]

    tc = pl.traj_tree_opt.config
    tcfg = dict(dt=tc.dt, state_size=tc.state_size, action_size=tc.action_size,
                w_opt_cfg={k: (np.asarray(v).copy() if isinstance(v, np.ndarray) else v) for k, v in tc.w_opt_cfg.items()},
                opt_cfg={k: (np.asarray(v).copy() if isinstance(v, np.ndarray) else v) for k, v in tc.opt_cfg.items()})
    sc = gen.config
    scfg = dict(max_depth=sc.max_depth, tar_dist_thres=sc.tar_dist_thres, tar_time_ahead=sc.tar_time_ahead)
    sim_time, step = 0.0, 0.02
    while sim_time < horizon:
        clock[0] = sim_time
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
                        ok, res = a.plan()                   # MINDAgent.plan (agent.py:324-327) -> the wrapped pl.plan below
                        assert ok, "plan failed at t=%.2f" % sim_time
            else:
                a.step()
            a.update_state(step)
        sim_time += step
    os.chdir(cwd)
    return dict(demo=demo, seq_id=seq, horizon=horizon, every=every, n_plan_calls=count[0], traj_cfg=tcfg, scen_cfg=scfg, records=records,
                host="build container: %d cores, torch %s CPU" % (os.cpu_count(), torch.__version__))


def annotate(out):
    """Distance of every keep / merge decision from its pi/6 threshold (scenario_tree.py:396-410) per recorded call, from the
    CPU oracle tree on the recorded inputs: the reference re-derives sub-metre displacements from global coordinates of
    several km in fp32 (conditioning note), so a decision within a few hundredths of a radian of the threshold
    has no well-defined outcome -- even this fp32 restatement of the same formulas lands on the other side for some
    recorded calls.  The replay test holds only the decisions outside that band to node-for-node identity."""
    from mind_b200.integration.replay import PlanReplayer
    from oracle.tree_oracle import TreeOracle
    from oracle.scene_pred_oracle import ScenePredOracle
    from types import SimpleNamespace
    sd = torch.load(os.path.join(OUT_DIR, "weights_20240121-172745.pt"), map_location="cpu")
    orc = ScenePredOracle(sd)

    class Net:
        def pre_process(self, d):
            return tuple(d[k] for k in ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"])

        def __call__(self, x):
            return orc(x)
    for r in out["records"]:
        t = TreeOracle(Net(), 50, 50, SimpleNamespace(**out["scen_cfg"]))
        t.reset()
        t.set_target_lane(r["lane"], r["info"])
        t.lane_graph = copy.deepcopy(r["graph"])
        trees = t.rollout(PlanReplayer.scene_dict(r))
        r["merge_margins"] = [(int(d), float(m)) for d, _, m in t.merge_margins]
        r["oracle_same_trees"] = [sorted(x.nodes) for x in trees] == [sorted(x) for x in r["scen_trees"]]
        print("  plan %d: closest merge decision %.4f rad from its threshold, oracle tree %s" %
              (r["plan_index"], min((abs(m) for _, m in r["merge_margins"]), default=9.9), "same" if r["oracle_same_trees"] else "differs"), flush=True)
    return out


def save(out, demo):
    buf = io.BytesIO()
    torch.save(out, buf)
    path = os.path.join(OUT_DIR, "plan_calls_%s.pt.xz" % demo)
    with open(path, "wb") as f:
        f.write(lzma.compress(buf.getvalue(), preset=6))
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("demos", nargs="*", default=["demo_1", "demo_2", "demo_3", "demo_4"])
    ap.add_argument("--horizon", type=float, default=10.0)
    ap.add_argument("--every", type=int, default=8)
    ap.add_argument("--annotate-only", action="store_true", help="add the oracle's decision margins to existing fixture files")
    args = ap.parse_args()
    for demo in args.demos:
        if args.annotate_only:
            from mind_b200.integration.replay import load_records
            out = load_records(os.path.join(OUT_DIR, "plan_calls_%s.pt.xz" % demo))
        else:
            out = record(demo, args.horizon, args.every)
        out = annotate(out)
        path = save(out, demo)
        print("%s: %d of %d plan calls recorded -> %s (%.1f MB)" % (demo, len(out["records"]), out["n_plan_calls"], path, os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
