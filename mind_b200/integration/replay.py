"""Replay of recorded MINDPlanner.plan calls on the product's planner stack (BASELINE.json configs[4] on the GPU box).

The closed loop itself needs MIND's simulator, agents and map classes, which stay the reference's own Python and cannot
travel to the GPU box.  oracle/record_plan_calls.py therefore records, while the UNMODIFIED reference drives a demo closed
loop on the CPU, what each plan call was given and what it answered; `PlanReplayer.plan(record)` is MINDPlanner.plan
(reference planners/mind/planner.py:104-145, :171-200) re-stated on this library's pieces:

    scenario trees     ScenarioTreeGeneratorB200.rollout on the CUDA predictor          (:107-113)
    trajectory trees   TrajectoryTreeOptimizerB200: warm start + full solve per tree   (:171-175)
                       = mind_cost_fields on the GPU + mind_ilqr_tree_solve
    selection          evaluate_traj_tree, minimum cost                                 (:133-145, :177-196)

so that the control it returns can be compared with the control the reference returned for the same inputs, and timed.
The scene dict enters where process_data's output would (the front end is host code, pinned separately on the CPU:
tests/test_front_end_cpu.py).
"""
import copy
import io
import lzma
import os
import time
from types import SimpleNamespace

import numpy as np
import torch

from .. import plumbing as P


def load_records(path):
    """tests/golden/plan_calls_<demo>.pt.xz -> dict (see oracle/record_plan_calls.py)"""
    with open(path, "rb") as f:
        return torch.load(io.BytesIO(lzma.decompress(f.read())), weights_only=False)


def dist_to_polyline(point, polyline):
    """distance of `point` from its projection on `polyline` (planner.py:198-203 with common/geometry.py:81-100)"""
    px, py = point
    sx, sy = polyline[:-1].T
    ex, ey = polyline[1:].T
    dx, dy = ex - sx, ey - sy
    t = np.clip(((px - sx) * dx + (py - sy) * dy) / (dx ** 2 + dy ** 2), 0, 1)
    nx, ny = sx + t * dx, sy + t * dy
    d = np.sqrt((px - nx) ** 2 + (py - ny) ** 2)
    i = int(np.argmin(d))
    return float(np.linalg.norm(np.array([nx[i], ny[i]]) - np.asarray(point)))


def evaluate_traj_tree(traj_tree, target_velocity, target_lane):
    """planner.py:177-196"""
    comfort = efficiency = target = 0.0
    for node in traj_tree.nodes.values():
        state, ctrl = node.data[0], node.data[1]
        comfort += 0.1 * ctrl[0] ** 2 + 5.0 * ctrl[1] ** 2
        efficiency += 0.01 * (target_velocity - state[2]) ** 2
        target += 0.01 * dist_to_polyline(state[:2], target_lane)
    return (comfort + efficiency + target) / len(traj_tree.nodes)


class PlanReplayer:
    def __init__(self, device, network, traj_cfg: dict, scen_cfg: dict):
        from ..scenario_tree import ScenarioTreeGeneratorB200
        from ..traj_opt import TrajectoryTreeOptimizerB200
        self.device = torch.device(device)
        self.gen = ScenarioTreeGeneratorB200(self.device, network, 50, 50, SimpleNamespace(**scen_cfg))
        self.opt = TrajectoryTreeOptimizerB200(SimpleNamespace(**traj_cfg), device=self.device)

    @staticmethod
    def scene_dict(rec):
        """the collated dict as process_data returned it: the recorder dropped the dense RPE, which is get_rpe of the
        anchors (scenario_tree.py:176-186, utils.py:193-212)"""
        data = copy.deepcopy(rec["data"])
        tj, graph = data["TRAJS"][0], data["LANE_GRAPH"][0]
        ctrs = torch.cat([tj["TRAJS_CTRS"], graph["lane_ctrs"]], 0)
        vecs = torch.cat([tj["TRAJS_VECS"], graph["lane_vecs"]], 0)
        data["RPE"] = [{"scene": P.pairwise_rpe(ctrs, vecs), "scene_mask": None}]
        return data

    def plan(self, rec):
        """-> dict(ctrl, best_idx, scen_trees, traj_trees, seconds = {scenario_tree, optimizer, total})"""
        data, graph = self.scene_dict(rec), copy.deepcopy(rec["graph"])        # fixture reconstruction: outside the timed call
        t0 = time.perf_counter()
        gen = self.gen
        gen.reset()                                                            # planner.py:107
        gen.set_target_lane(rec["lane"], rec["info"])                          # :109-111
        gen.lane_graph = graph
        scen_trees = gen.rollout(data)                                         # :113 (behind process_data)
        t1 = time.perf_counter()
        traj_trees = []
        for st in scen_trees:                                                  # :120-123, get_traj_tree :171-175
            self.opt.init_warm_start_cost_tree(st, rec["state"], rec["ctrl"], rec["gt_tgt_lane"], rec["target_velocity"])
            _, us = self.opt.warm_start_solve()
            self.opt.init_cost_tree(st, rec["state"], rec["ctrl"], rec["gt_tgt_lane"], rec["target_velocity"])
            traj_trees.append(self.opt.solve(us))
        best, min_cost = None, np.inf                                          # :133-140
        for i, tt in enumerate(traj_trees):
            c = evaluate_traj_tree(tt, rec["target_velocity"], rec["lcl_target_lane"])
            if c < min_cost:
                min_cost, best = c, i
        tt = traj_trees[best]
        nxt = tt.get_node(tt.get_root().children_keys[0])                      # :142-144
        ctrl = np.asarray(nxt.data[0][-2:], dtype=np.float64)
        t2 = time.perf_counter()
        return dict(ctrl=ctrl, best_idx=best, scen_trees=scen_trees, traj_trees=traj_trees,
                    seconds=dict(scenario_tree=t1 - t0, optimizer=t2 - t1, total=t2 - t0))


def replay_file(path, device, network, warmup=1):
    """all records of one demo -> list of per-call results with the reference's answers next to them"""
    rec = load_records(path)
    rp = PlanReplayer(device, network, rec["traj_cfg"], rec["scen_cfg"])
    out = []
    for i, r in enumerate(rec["records"]):
        for _ in range(warmup if i == 0 else 0):
            rp.plan(r)
        res = rp.plan(r)
        ref_keys = [sorted(t) for t in r["scen_trees"]]
        got_keys = [sorted(t.nodes) for t in res["scen_trees"]]
        out.append(dict(plan_index=r["plan_index"], sim_time=r["sim_time"], ctrl=res["ctrl"], ref_ctrl=r["ret_ctrl"],
                        same_trees=ref_keys == got_keys, got_keys=got_keys, ref_keys=ref_keys, best_idx=res["best_idx"],
                        ref_best=r["best_candidates"], seconds=res["seconds"], ref_seconds=r["cpu_reference_s"],
                        n_trees=len(got_keys), merge_margins=r.get("merge_margins")))
    return rec, out


FRAGILE_RAD = 0.05      # a keep / merge decision closer than this to pi/6 can fall either way (see oracle/record_plan_calls.py)


def fragile_depth(result):
    """first tree depth at which the recorded call has a decision inside the noise band, or None"""
    mm = result.get("merge_margins")
    if not mm:
        return None
    d = [int(dep) for dep, m in mm if abs(m) < FRAGILE_RAD]
    return min(d) if d else None


def keys_above(keys, depth):
    """node keys '{depth}_{scene}_{mode}' created before `depth`, over all trees of a call"""
    return sorted(k for t in keys for k in t if int(k.split("_")[0]) < depth)
