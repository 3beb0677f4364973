"""Run MIND's own closed-loop simulator (run_sim.py -> simulator.Simulator, unmodified) with pieces of this library
dropped in, BASELINE.json configs[4]:

  python -m mind_b200.integration.run_sim --mind-root /path/to/MIND --demo demo_2 \\
         [--predictor b200|reference] [--tree b200|reference] [--optimizer b200|reference] [--render] [--horizon STEPS]

* `--predictor b200`  is configuration only: a planner JSON whose `network_config` names
  `mind_b200.integration.net_cfg_b200` (the reference resolves the class from that string, planner.py:42-49).
* `--tree b200` / `--optimizer b200` replace two attributes of the constructed planner (`scen_tree_gen`,
  `traj_tree_opt`, planner.py:51-57) with `ScenarioTreeGeneratorB200` / `TrajectoryTreeOptimizerB200`: same call surface.
* av2 / shapely / Theano / matplotlib are taken from `mind_b200.compat` when the real packages are absent
  (without matplotlib the run is forced to `"render": false`).
The B200 pieces need a CUDA device; `--predictor reference --tree reference --optimizer reference` is the reference arm.
"""
import argparse
import json
import os
import sys
import tempfile
import time


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--mind-root", default=os.environ.get("MIND_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--demo", default="demo_2")
    ap.add_argument("--predictor", choices=["b200", "reference"], default="b200")
    ap.add_argument("--tree", choices=["b200", "reference"], default="b200")
    ap.add_argument("--optimizer", choices=["b200", "reference"], default="b200")
    ap.add_argument("--render", action="store_true")
    ap.add_argument("--horizon", type=int, default=None, help="simulation steps of 0.02 s (reference: 500)")
    args = ap.parse_args(argv)
    root = os.path.abspath(args.mind_root)
    if not os.path.isdir(os.path.join(root, "planners", "mind")):
        raise SystemExit("no MIND checkout at %s (--mind-root)" % root)
    from mind_b200 import compat
    used = compat.install()
    print("third-party packages:", used)
    sys.path.insert(0, root)
    os.chdir(root)                                              # the reference uses paths relative to its root
    sim_cfg = json.load(open(os.path.join(root, "configs", args.demo + ".json")))
    if not args.render or used.get("matplotlib", "").startswith("stub"):
        sim_cfg["render"] = False
    tmp = tempfile.mkdtemp(prefix="mind_b200_")
    for agent in sim_cfg["cl_agents"]:
        pc = json.load(open(os.path.join(root, agent["planner_config"])))
        if args.predictor == "b200":
            pc["network_config"] = "mind_b200.integration.net_cfg_b200"
            pc["use_cuda"] = True
        path = os.path.join(tmp, "planner_%s.json" % agent["id"])
        json.dump(pc, open(path, "w"))
        agent["planner_config"] = path
    sim_path = os.path.join(tmp, "sim.json")
    json.dump(sim_cfg, open(sim_path, "w"))
    from simulator import Simulator
    sim = Simulator(sim_path)
    if args.horizon is not None:
        sim.sim_horizon = args.horizon
    sim.init_sim()
    for a in sim.agents:
        pl = getattr(a, "planner", None)
        if pl is None:
            continue
        if args.tree == "b200":
            from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
            old = pl.scen_tree_gen
            pl.scen_tree_gen = ScenarioTreeGeneratorB200(pl.device, pl.network, pl.obs_len, pl.plan_len, old.config)
        if args.optimizer == "b200":
            from mind_b200.traj_opt import TrajectoryTreeOptimizerB200
            pl.traj_tree_opt = TrajectoryTreeOptimizerB200(pl.traj_tree_opt.config, device=pl.device)
    t0 = time.perf_counter()
    sim.run_sim()
    wall = time.perf_counter() - t0
    sim.render_video()
    ego = [a for a in sim.agents if getattr(a, "planner", None) is not None][0]
    print(json.dumps({"demo": args.demo, "predictor": args.predictor, "tree": args.tree, "optimizer": args.optimizer,
                      "sim_steps": len(sim.frames), "wall_s": round(wall, 2), "ego_final_state": [round(float(v), 4) for v in ego.state]}))


if __name__ == "__main__":
    main()
