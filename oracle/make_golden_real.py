"""Golden vectors on the REAL Argoverse-2 demo scenes (BASELINE.json configs 1 and 3), produced by the UNMODIFIED
reference: SemanticMap, ArgoAgentLoader, the agents' replay loop of simulator.py:52-107, MINDPlanner.update_observation /
resample_target_lane, ScenarioTreeGenerator.process_data, ScenePredNet and branch_aime.      python -m oracle.make_golden_real

Test infrastructure, build container only (needs /root/reference).  av2 / shapely come from mind_b200.compat (the real
packages are not in the image: that boundary is parity-unpinned).  The closed-loop agent is never enabled here, so the
ego vehicle stays on its recorded trajectory (open loop; oracle/run_closed_loop.py drives it closed loop).  At sim time T
the script does
exactly what MINDPlanner.plan does up to the scenario trees (planner.py:104-112).

Writes tests/golden/real_<demo>.pt per demo scene:
  data     collated scene dict returned by process_data (network inputs + tree bookkeeping), CPU tensors
  lane, info, graph   target lane / info arrays passed to set_target_lane, generator.lane_graph
  cls, reg, vel       reference network outputs on `data`
  tree     {key: (parent, prob, trajs, covs, tgt)} of branch_aime, levels = net batch size per depth level
  level_inputs   the collated network inputs of every depth level >= 1 as the reference built them (update_obser)
"""
import copy
import json
import os
import sys
import types

import numpy as np
import torch

from oracle import ref_loader
from oracle.tree_oracle import flatten_trees

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "tests", "golden")
DEMOS = {"demo_1": 5.0, "demo_2": 5.0, "demo_3": 5.0, "demo_4": 5.0}      # sim time of the captured plan call
# the closed-loop agent's FIRST plan call (enable_timestep 4.0 s): only 41 observed steps, every track is front-padded
EXTRA = {"demo_3_t4": ("demo_3", 4.0)}


def load_reference_sim():
    from mind_b200 import compat
    compat.install()                                          # av2 / shapely / theano stand-ins where the real ones are absent
    if ref_loader.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REF_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.semantic_map = importlib.import_module("common.semantic_map")
    ns.planner = importlib.import_module("planners.mind.planner")
    ns.agent = importlib.import_module("agent")
    ns.loader = importlib.import_module("loader")
    ns.utils = importlib.import_module("planners.mind.utils")
    return ns


def run_until(ns, cfg, t_plan):
    """simulator.py:52-107 without rendering and without enabling the closed-loop agent; returns the MIND agent."""
    from pathlib import Path
    seq = cfg["seq_id"]
    seq_path = os.path.join(ref_loader.REF_ROOT, "data", seq)
    smp = ns.semantic_map.SemanticMap()
    smp.load_from_argo2(Path(os.path.join(seq_path, "log_map_archive_%s.json" % seq)))
    cwd = os.getcwd()
    os.chdir(ref_loader.REF_ROOT)                              # planner configs use paths relative to the reference root
    try:
        agents = ns.loader.ArgoAgentLoader(Path(os.path.join(seq_path, "scenario_%s.parquet" % seq))).load_agents(smp, cfg["cl_agents"])
    finally:
        os.chdir(cwd)
    sim_time, step = 0.0, 0.02
    ego = [a for a in agents if isinstance(a, ns.agent.CustomizedAgent)][0]
    while True:
        obs = [a.observe() for a in agents
               if (isinstance(a, ns.agent.NonReactiveAgent) and a.is_valid()) or isinstance(a, ns.agent.CustomizedAgent)]
        for a in agents:
            if isinstance(a, ns.agent.CustomizedAgent):
                rec_tri, pl_tri = a.check_trigger(sim_time)    # never enabled: replay + observe
                if rec_tri:
                    a.step()
                if pl_tri:
                    a.update_observation(obs)
                    if sim_time >= t_plan - 1e-9:
                        return a, smp
            else:
                a.step()
            a.update_state(step)
        sim_time += step


def capture(ns, name, t_plan):
    cfg = json.load(open(os.path.join(ref_loader.REF_ROOT, "configs", name + ".json")))
    ego, smp = run_until(ns, cfg, t_plan)
    pl = ego.planner
    gen = pl.scen_tree_gen
    gen.reset()                                                # planner.py:106-112
    lane, info = pl.resample_target_lane(ego.lcl_smp)
    gen.set_target_lane(lane, info)
    with torch.no_grad():
        data = gen.process_data(ego.lcl_smp, pl.agent_obs)
        graph = copy.deepcopy(gen.lane_graph)
        res_cls, res_reg, res_aux = gen.network(gen.network.pre_process(data))
        # branch_aime (scenario_tree.py:38-58) with the level sizes recorded
        gen.reset()
        gen.set_target_lane(lane, info)
        data2 = gen.process_data(ego.lcl_smp, pl.agent_obs)
        gen.init_scenario_tree(data2)
        levels, level_inputs = [1], []
        nodes = gen.get_branch_set()
        while nodes:
            batch = ns.utils.collate_fn([n.data.obs_data for n in nodes])
            levels.append(len(nodes))
            level_inputs.append({k: copy.deepcopy(batch[k]) for k in
                                 ("ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE")})
            gen.create_nodes(gen.prune_merge(batch, gen.predict_scenes(batch)))
            gen.decide_branch()
            nodes = gen.get_branch_set()
        flat = flatten_trees(gen.get_scenario_tree())
    out = dict(name=name, t_plan=t_plan, seq_id=cfg["seq_id"], data=data, lane=np.asarray(lane), info=[np.asarray(x) for x in info],
               graph=graph, cls=[c.clone() for c in res_cls], reg=[r.clone() for r in res_reg], vel=[a[0].clone() for a in res_aux],
               tree=flat, levels=levels, level_inputs=level_inputs)
    na, nl = data["TRAJS"][0]["TRAJS_POS_OBS"].shape[0], graph["lane_ctrs"].shape[0]
    print("%s t=%.1f: %d actors, %d lane polylines, target lane %d pts, levels %s, %d tree nodes" %
          (name, t_plan, na, nl, len(lane), levels, len(flat)))
    return out


def main():
    ns = load_reference_sim()
    for name, t_plan in DEMOS.items():
        torch.save(capture(ns, name, t_plan), os.path.join(OUT_DIR, "real_%s.pt" % name))
    for tag, (name, t_plan) in EXTRA.items():
        torch.save(capture(ns, name, t_plan), os.path.join(OUT_DIR, "real_%s.pt" % tag))


if __name__ == "__main__":
    main()
