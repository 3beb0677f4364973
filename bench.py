#!/usr/bin/env python
"""bench.py -- scene-predictions/sec (K=6 modes) of the MIND scenario-prediction hot path.

Workload (BASELINE.json configs[1]): batch=256 synthetic scenes per GPU, 32 actors x 128 lane
polylines (N=161 tokens), K=6 modes, seeded inputs (mind_b200/synth.py), weights = the reference's
shipped checkpoint (tests/golden/weights_*.pt).  One "step" = one batched forward of the whole
batch: encoders + 6 rela-fusion layers + decoder, outputs in the reference layout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

N>1 is launched by torchrun (one rank per GPU); every rank predicts its own 256 scenes (weak
scaling) and the step ends with one NCCL all-gather of the decoded trajectories (configs[3]).
`--impl reference` times the reference algorithm's CPU port (oracle/) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scene-predictions/sec (K=6 modes)"
UNIT = "scenes/s"
NA, NL, D = 32, 128, 128
N_TOK = NA + NL + 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def load_weights():
    return torch.load(os.path.join(ROOT, "tests", "golden", "weights_20240121-172745.pt"), map_location="cpu")


def make_batch(batch, seed0):
    from mind_b200 import synth
    return synth.batch_s2(batch=batch, n_actor=NA, n_lane=NL, seed0=seed0)


def make_batch_dict(batch, seed0):
    """The collated dict network.pre_process receives (scenario_tree.py:69-70): the 7 network inputs plus the per-scene
    anchors the dense RPE was built from (TRAJS[b]['TRAJS_CTRS'/'TRAJS_VECS'], LANE_GRAPH[b]['lane_ctrs'/'lane_vecs'],
    scenario_tree.py:160-186, utils.py:468-483), as collate_fn leaves them (lists over scenes)."""
    from mind_b200 import synth
    scenes = [synth.scene_s1(seed0 + b, NA, NL, with_geom=True) for b in range(batch)]
    keys = ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"]
    d = dict(zip(keys, synth.batch_from_scenes(scenes)))
    d["TRAJS"] = [{"TRAJS_CTRS": s["ctrs"][:NA].contiguous(), "TRAJS_VECS": s["vecs"][:NA].contiguous()} for s in scenes]
    d["LANE_GRAPH"] = [{"lane_ctrs": s["ctrs"][NA:].contiguous(), "lane_vecs": s["vecs"][NA:].contiguous()} for s in scenes]
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def fusion_flops_per_scene(update_edge=True):
    """canonical (concat-split) algorithmic FLOPs of the N^2 part of one layer, SURVEY.md 8d:
    2 N^2 D^2 (3+u) + 4 N^2 D   (W_e, W_pe, K, V on N^2 rows; QK^T and PV)."""
    u = 1 if update_edge else 0
    return 2.0 * N_TOK * N_TOK * D * D * (3 + u) + 4.0 * N_TOK * N_TOK * D


def fusion_bytes_per_scene(update_edge=True):
    """fp16 edge stream: read N^2*128*2 B, write the same when the layer updates the edge."""
    return N_TOK * N_TOK * D * 2.0 * (2 if update_edge else 1)


def cpu_layout():
    """(processes, threads per process) for the CPU port.  torch's intra-op pool degrades badly past ~16 threads on this
    workload (measured on the 128-core GPU-box host: 128 threads -> 0.03 scenes/s, 32 s/scene), so the host cores are
    used as several 16-thread processes, each predicting its own scenes."""
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, int(os.environ.get("MIND_CPU_THREADS", "16"))))
    procs = max(1, min(cores // threads, int(os.environ.get("MIND_CPU_PROCS", "8"))))
    return procs, threads


def _cpu_worker(idx, threads, chunk, seed0, seconds, warm, q, go):
    torch.set_num_threads(threads)
    from oracle.scene_pred_oracle import ScenePredOracle
    orc = ScenePredOracle(load_weights())
    data = make_batch(chunk, seed0 + 100 * idx)
    for _ in range(warm):
        orc(data)
    q.put(("ready", idx))
    go.wait()
    n, t0 = 0, time.perf_counter()
    while True:
        orc(data)
        n += chunk
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    q.put(("done", idx, n, el))


def cpu_port_rate(sd, seconds_target=10.0, chunk=2, seed0=5000):
    """Reference algorithm on the host cores: the oracle port (torch CPU fp32) in `procs` processes of `threads` threads,
    bounded sample of the same workload (S2 scenes 32x128).  Returns (scenes/s over all processes, cores used, sample)."""
    import multiprocessing as mp
    procs, threads = cpu_layout()
    ctx = mp.get_context("spawn")
    q, go = ctx.Queue(), ctx.Event()
    ps = [ctx.Process(target=_cpu_worker, args=(i, threads, chunk, seed0, seconds_target, 1, q, go)) for i in range(procs)]
    for p_ in ps:
        p_.start()
    for _ in ps:
        q.get(timeout=600)
    go.set()
    res = [q.get(timeout=600) for _ in ps]
    for p_ in ps:
        p_.join(timeout=60)
    n = sum(r[2] for r in res)
    el = max(r[3] for r in res)
    return n / el, procs * threads, ("%d S2 scenes (32 actors x 128 lanes) in %.1f s, %d process(es) x %d threads, batches of %d (host has %d cores)"
                                     % (n, el, procs, threads, chunk, os.cpu_count() or 1))


def gpu_eager_rate(sd, dev, batches=(8, 64)):
    """Second baseline: the reference network's arithmetic under torch EAGER on this same GPU (cuBLAS /
    ATen kernels, fp32, the reference's Python loop over scenes, network.py:318,497).  The reference module itself cannot
    travel to the box, so this is the oracle's functional restatement of it with its tensors on the device (kind "port";
    same op sequence as the nn.Module: Linear / LayerNorm / softmax / einsum per scene).  Never fatal."""
    out = {"kind": "port (oracle/scene_pred_oracle.py on cuda, torch eager fp32, TF32 off)", "unit": UNIT}
    try:
        from oracle.scene_pred_oracle import ScenePredOracle
        torch.backends.cuda.matmul.allow_tf32 = False
        orc = ScenePredOracle(sd, device=dev)
        for nb in batches:
            data = make_batch(nb, 7000)
            data = (data[0].to(dev), data[1], data[2].to(dev), data[3], [{"scene": r["scene"].to(dev), "scene_mask": None} for r in data[4]],
                    data[5].to(dev), data[6].to(dev))
            orc(data)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 if nb <= 8 else 2
            e0.record()
            for _ in range(reps):
                orc(data)
            e1.record()
            torch.cuda.synchronize()
            out["scenes_per_s_B%d" % nb] = nb * reps / (e0.elapsed_time(e1) * 1e-3)
    except Exception as e:
        out["error"] = repr(e)[:300]
    return out


class _TreeCfg:      # planners/mind/configs/planning/demo_1.py:3-10
    max_depth = 5
    tar_dist_thres = 10.0
    tar_time_ahead = 5.0


def bench_tree(net, dev, reps=5):
    """tree-rollout ms/scene (BASELINE.json configs[2] stand-in): S3 kinematic scene (8 actors, 60 lane
    polylines); (i) the natural AIME tree, (ii) forced-full depth 4 x branch 6 (level batches 1/6/36/216,
    259 scene predictions, 1296 leaves).  Wall time from the collated root scene to the packed trees."""
    import copy
    from mind_b200 import synth
    from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
    args = dict(x0=(100, 108, 92, 120, 112, 96, 130, 85), y0=(0, 3.5, -3.5, 0, 3.5, 3.5, -3.5, 0), v=(5, 9, 3, 8, 2, 10, 6, 12))
    out = {"tiers": "natural trees: tensor-core mode as shipped (these scenes have < 128 tokens -> exact tier, bit-exact branch selection); "
                    "forced_full: every keep / branch decision is forced, so nothing depends on the predictions' last bits -> fused "
                    "fp16-operand tier (tc_min_tokens = 0); forced_full_exact_tier: the same tree in the exact tier"}
    for name, ff in (("natural", None), ("forced_full", (10, 20, 30)), ("forced_full_exact_tier", (10, 20, 30))):
        gen = ScenarioTreeGeneratorB200(dev, net, 50, 50, _TreeCfg())
        gen.force_full = ff
        net.set_option("tc_min_tokens", 0 if name == "forced_full" else 128)
        times = []
        for r in range((reps if name != "forced_full_exact_tier" else 1) + 2):
            data, lane, info, graph = synth.scene_s3(**args)
            gen.reset(); gen.set_target_lane(lane, info); gen.lane_graph = copy.deepcopy(graph)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            trees = gen.rollout(data)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        out[name] = {"ms_per_tree": statistics.median(times[2:]), "level_batches": list(gen.net_batches),
                     "nodes": gen.tree.size(), "trees": len(trees),
                     "host_phase_ms_last": {k: round(v * 1e3, 3) for k, v in gen.timing.items()}}
    # BASELINE.json configs[2] itself: the demo_2 Argoverse-2 scene (45 actors x 37 lane polylines at sim time 5.0 s) as
    # the unmodified reference's process_data built it (tests/golden/real_demo_2.pt, oracle/make_golden_real.py)
    fx = os.path.join(ROOT, "tests", "golden", "real_demo_2.pt")
    if os.path.exists(fx):
        gold = torch.load(fx, weights_only=False)
        for name, ff in (("demo_2_natural", None), ("demo_2_forced_full", (10, 20, 30))):
            gen = ScenarioTreeGeneratorB200(dev, net, 50, 50, _TreeCfg())
            gen.force_full = ff
            net.set_option("tc_min_tokens", 0 if ff else 128)
            times = []
            try:
                for r in range(reps + 2):
                    data = copy.deepcopy(gold["data"])
                    gen.reset(); gen.set_target_lane(gold["lane"], gold["info"]); gen.lane_graph = copy.deepcopy(gold["graph"])
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    trees = gen.rollout(data)
                    torch.cuda.synchronize()
                    times.append((time.perf_counter() - t0) * 1e3)
            except Exception as e:                                  # reported, never fatal for the headline line
                out[name] = {"error": repr(e)[:300]}
                continue
            out[name] = {"ms_per_tree": statistics.median(times[2:]), "level_batches": list(gen.net_batches),
                         "nodes": gen.tree.size(), "trees": len(trees), "actors": int(gold["data"]["ACTORS"].shape[0]),
                         "lane_polylines": int(gold["data"]["LANES"].shape[0]), "input": "host tensors (collated scene dict on the CPU)",
                         "host_phase_ms_last": {k: round(v * 1e3, 3) for k, v in gen.timing.items()}}
    net.set_option("tc_min_tokens", 128)
    return out


def bench_cost_fields(dev, reps=5):
    """SURVEY.md 8f-3 (the step right after the path): cost fields of the trajectory-tree optimiser for the demo_2 scenario
    trees at the reference's 256 x 256 x 0.4 m grid (planners/mind/configs/planning/demo_*.py:73-81).  GPU: wall time of
    mind_b200.cost_field.cost_fields per tree (tables H2D + mind_cost_fields + fields D2H) and CUDA-event time of the two
    kernels alone; CPU: the oracle's numpy restatement of trajectory_tree.py:58-124 on one tree.  Never fatal."""
    try:
        import ctypes as C
        import numpy as np
        from mind_b200 import cost_field as CF, lib as L
        from oracle import cost_field_oracle as O
        fx = os.path.join(ROOT, "tests", "golden", "real_demo_2.pt")
        flat = torch.load(fx, weights_only=False)
        lane, flat = np.asarray(flat["lane"], dtype=np.float64), flat["tree"]
        kids = {}
        for k, v in flat.items():
            kids.setdefault(v[0], []).append(k)

        class N:
            def __init__(self, k):
                self.key, self.parent_key, self.children_keys = k, flat[k][0], sorted(kids.get(k, []))
                self.data = [flat[k][1], flat[k][2], flat[k][3], None]

        class T:
            def __init__(self, root):
                self.root, self.nodes = root, {}
                todo = [root]
                while todo:
                    k = todo.pop()
                    self.nodes[k] = N(k)
                    todo += self.nodes[k].children_keys
            get_root = lambda self: self.nodes[self.root]
            get_node = lambda self, k: self.nodes[k]
        trees = [T(k) for k in sorted(k for k, v in flat.items() if v[0] is None)]
        cfg = dict(w_tgt=1.0, w_ego=1.0, w_ego_cov_offset=1.0, w_exo=10.0, w_exo_cov_offset=2.5, w_exo_cost_offset=10.0,
                   smooth_grid_size=(256, 256), smooth_grid_res=0.4)
        ego = flat[trees[0].root][2][0, 0]
        x0 = np.array([ego[0], ego[1], 6.0, 0.1, 0.2, 0.01])
        wall, nodes = [], 0
        for r in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nodes = 0
            for t in trees:
                nodes += len(CF.cost_fields(t, x0, lane, cfg, dev, warm=False)["links"])
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3 / len(trees))
        # kernels alone: same tables, device-resident, CUDA events on the launching stream
        t = trees[0]
        coef, mean, rad, links, _ = CF.node_tables(t, cfg, False)
        off, xs, ys = CF.grid_frame(x0, cfg["smooth_grid_size"], cfg["smooth_grid_res"])
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        bufs = [d(xs), d(ys), d(lane), d(coef), d(mean), d(rad), torch.empty(256, 256, dtype=torch.float64, device=dev),
                torch.empty(len(coef), 256, 256, dtype=torch.float64, device=dev)]
        a = L.MindCostFields()
        a.gx, a.gy, a.xs, a.ys, a.n_lane_pts, a.lane = 256, 256, bufs[0].data_ptr(), bufs[1].data_ptr(), len(lane), bufs[2].data_ptr()
        a.n_nodes, a.n_actor, a.coef_tgt, a.mean, a.radius = len(coef), mean.shape[1], bufs[3].data_ptr(), bufs[4].data_ptr(), bufs[5].data_ptr()
        a.w_ego, a.w_exo, a.exo_cost_offset, a.quad, a.fields = 1.0, 10.0, 10.0, bufs[6].data_ptr(), bufs[7].data_ptr()
        lib = L.load()
        st = torch.cuda.current_stream(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for r in range(3):
            lib.mind_cost_fields(C.byref(a), C.c_void_p(st.cuda_stream))
        ev[0].record(st)
        for r in range(10):
            lib.mind_cost_fields(C.byref(a), C.c_void_p(st.cuda_stream))
        ev[1].record(st)
        torch.cuda.synchronize()
        k_ms = ev[0].elapsed_time(ev[1]) / 10
        by = len(coef) * 256 * 256 * 8.0
        # CPU: numpy restatement on the same tree (the reference's own loop structure), one pass
        onodes = {k: (n.parent_key, n.data[0], n.data[1], n.data[2], n.children_keys) for k, n in t.nodes.items()}
        t0 = time.perf_counter()
        O.cost_fields(onodes, t.root, x0, lane, cfg, warm=False)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        # the whole optimiser step per tree: fields on the GPU + native tree iLQR (warm start from zero controls, then the full solve)
        opt_ms = None
        try:
            from mind_b200.traj_opt import solve_tree
            ocfg = dict(cfg, w_des_state=np.diag([0, 0, 0.1, 0, 1.0, 10.0]), w_state_con=np.diag([0, 0, 50.0, 0, 50.0, 500.0]),
                        state_upper_bound=np.array([1e5, 1e5, 8.0, 10.0, 4.0, 0.2]), state_lower_bound=np.array([-1e5, -1e5, 0.0, -10.0, -6.0, -0.2]),
                        w_ctrl=5.0 * np.eye(2))
            tt = []
            for r in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for tr in trees:
                    _, us_w, _ = solve_tree(tr, x0, lane, 8.0, ocfg, 0.2, warm=True, device=dev)
                    solve_tree(tr, x0, lane, 8.0, ocfg, 0.2, us_init=us_w, warm=False, device=dev)
                tt.append((time.perf_counter() - t0) * 1e3 / len(trees))
            opt_ms = min(tt)
        except Exception as e:
            opt_ms = repr(e)[:200]
        return {"workload": "demo_2 scenario trees (%d trees, %d trajectory-tree nodes, 45 actors), 256x256 cells of 0.4 m, fp64" % (len(trees), nodes),
                "optimizer_ms_per_tree_gpu_fields_plus_native_ilqr": opt_ms,
                "gpu_ms_per_tree_incl_copies": statistics.median(wall[1:]), "kernels_ms_per_tree": k_ms, "nodes_in_timed_tree": len(coef),
                "field_bytes_per_tree": by, "achieved_write_gbs": by / (k_ms * 1e-3) / 1e9,
                "cpu_numpy_ms_per_tree": cpu_ms, "cpu_kind": "port (oracle/cost_field_oracle.py, 1 numpy thread)"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def bench_closed_loop(net, dev):
    """BASELINE.json configs[4]: plan calls recorded while the UNMODIFIED reference drove demo_1..4 closed loop on the CPU
    (oracle/record_plan_calls.py) replayed on the product's planner stack on this GPU (mind_b200/integration/replay.py:
    scenario tree on the CUDA predictor, GPU cost fields, native tree iLQR).  Wall time per plan call from the collated
    scene dict (host) to the returned control, next to the reference's time for the same call on the CPU it was recorded
    on, and the largest control difference.  Never fatal."""
    import glob
    import numpy as np
    out = {}
    try:
        from mind_b200.integration.replay import fragile_depth, replay_file
        for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "plan_calls_demo_*.pt.xz"))):
            rec, res = replay_file(path, dev, net)
            tot = [r["seconds"]["total"] for r in res]
            out[rec["demo"]] = {
                "plan_calls_replayed": len(res), "plan_calls_in_run": rec["n_plan_calls"],
                "ms_per_plan_call": 1e3 * statistics.median(tot),
                "scenario_tree_ms": 1e3 * statistics.median(r["seconds"]["scenario_tree"] for r in res),
                "optimizer_ms": 1e3 * statistics.median(r["seconds"]["optimizer"] for r in res),
                "trees_per_call": sum(r["n_trees"] for r in res) / len(res),
                "reference_cpu_ms_per_plan_call": 1e3 * statistics.median(r["ref_seconds"]["scenario_tree"] + r["ref_seconds"]["optimizer"] for r in res),
                "reference_cpu_host": rec["host"],
                "calls_with_reference_trees_node_for_node": sum(r["same_trees"] for r in res),
                "calls_with_every_decision_clear_of_its_threshold": sum(fragile_depth(r) is None for r in res),
                "same_chosen_tree_when_trees_agree": all(r["best_idx"] in r["ref_best"] for r in res if r["same_trees"]),
                "max_abs_ctrl_diff_when_trees_agree": [float(v) for v in np.max([np.abs(r["ctrl"] - r["ref_ctrl"]) for r in res if r["same_trees"]] or [[0.0, 0.0]], axis=0)],
                "max_abs_ctrl_diff_all_calls": [float(v) for v in np.max([np.abs(r["ctrl"] - r["ref_ctrl"]) for r in res], axis=0)]}
        out["note"] = ("replay of recorded closed-loop plan calls (inputs = what the reference's process_data produced at that call); "
                       "the front end (process_data, host code) is not inside these times")
    except Exception as e:
        out["error"] = repr(e)[:300]
    return out


def run_reference(args):
    """--impl reference: the reference algorithm's CPU port (oracle/) on all the host cores it can use (several 16-thread
    processes); each "step" is a bounded sample of the workload: every process predicts 4 of the 256 scenes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    procs, threads = cpu_layout()
    per_step = 4
    ctx = mp.get_context("spawn")
    q, go = ctx.Queue(), ctx.Event()
    ps = [ctx.Process(target=_ref_worker, args=(i, threads, per_step, args.warmup, args.steps, q, go)) for i in range(procs)]
    for p_ in ps:
        p_.start()
    for _ in ps:
        q.get(timeout=1200)
    go.set()
    res = [q.get(timeout=3000) for _ in ps]
    for p_ in ps:
        p_.join(timeout=60)
    el = max(r[1] for r in res)
    v = procs * per_step * args.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "S2: 32 actors x 128 lanes, K=6; CPU sample of %d scenes/step (%d processes x %d scenes)" % (procs * per_step, procs, per_step)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs * threads, "kind": "port",
                             "sample": "%d processes x %d threads, %d scenes/step each x %d steps, oracle port (torch CPU fp32); host has %d cores"
                                       % (procs, threads, per_step, args.steps, os.cpu_count() or 1)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _ref_worker(idx, threads, per_step, warmup, steps, q, go):
    torch.set_num_threads(threads)
    from oracle.scene_pred_oracle import ScenePredOracle
    orc = ScenePredOracle(load_weights())
    data = make_batch(per_step, 1000 + 10 * idx)
    for _ in range(max(1, warmup)):
        orc(data)
    q.put(("ready", idx))
    go.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        orc(data)
    q.put(("done", time.perf_counter() - t0))


def run_native(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mind_b200.predictor import ScenePredNetB200
    sd = load_weights()
    net = ScenePredNetB200(None, dev)
    net.load_state_dict(sd)
    net.set_precision(args.precision)
    B = args.batch
    keys = ["ACTORS", "ACTOR_IDCS", "LANES", "LANE_IDCS", "RPE", "TGT_NODES", "TGT_RPE"]
    host_full = make_batch_dict(B, 1000 + rank * B)
    host = tuple(host_full[k] for k in keys)

    def pin(x):
        if isinstance(x, torch.Tensor):
            return x.pin_memory()
        if isinstance(x, list):
            return [pin(v) for v in x]
        if isinstance(x, dict):
            return {k: pin(v) for k, v in x.items()}
        return x
    host_dict = {k: pin(v) for k, v in host_full.items()}
    data_dev = net.pre_process(host_dict)
    torch.cuda.synchronize()
    # the path's single collective: ONE all-gather of the buffer the decoder kernels wrote (cls | reg | vel), issued
    # asynchronously so that it runs under the next step's encoders; two gather buffers alternate
    from mind_b200.distributed import all_gather_packed
    gather_bufs, pending = [None, None], [None, None]
    step_no = [0]

    def step_device():
        out = net.forward_packed(data_dev)
        if world > 1:
            k = step_no[0] & 1
            step_no[0] += 1
            if pending[k] is not None:
                pending[k][0].wait()                      # the gather that last used this buffer
            gather_bufs[k], work = all_gather_packed(out[6], B, out[1].shape[0], out=gather_bufs[k], async_op=True)
            pending[k] = (work, out)                      # keeps the send buffer alive until the collective is done
        return out

    def drain():
        for k in (0, 1):
            if pending[k] is not None:
                pending[k][0].wait()
                pending[k] = None

    for _ in range(args.warmup):
        step_device()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    net.profile(True)
    net.profile_read()
    l0 = net.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step_device()
    drain()                                                # every step's all-gather is inside the timed region
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = net.launch_count() - l0
    prof = net.profile_read()
    net.profile(False)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    if args.kernel_only:      # profiling runs (ncu): device-resident leg only
        if rank == 0:
            print(json.dumps({"kernel_only": True, "value": value, "ms_per_step": ms / args.steps, "gpu_launches": launches,
                              "stage_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items())}}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end to end through the reference-facing plugin call, host buffers in, host result out ----
    # Serving loop with two-deep pipelining: while step i computes, the copy stream uploads step i+1's
    # inputs (pre_process = the reference's own H2D point, network.py:597-606) and downloads step i-1's
    # results.  Every step's inputs start in pinned host memory and its cls/reg/vel end in host memory.
    copy_s = torch.cuda.Stream(dev)
    comp_s = torch.cuda.current_stream(dev)
    out_host = [None, None]
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_out = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    staged = [None, None]

    def upload(slot):
        with torch.cuda.stream(copy_s):
            staged[slot] = net.pre_process(host_dict)           # H2D of this step's 7 inputs (pinned -> device)
            ev_in[slot].record(copy_s)

        def mark(x):                                            # inputs are consumed on the compute stream
            if isinstance(x, torch.Tensor):
                if x.is_cuda:                                   # the index lists stay on the host
                    x.record_stream(comp_s)
            elif isinstance(x, (list, tuple)):
                for v in x:
                    mark(v)
            elif isinstance(x, dict):
                for v in x.values():
                    mark(v)
            elif hasattr(x, "ctrs") and hasattr(x, "vecs"):       # anchors uploaded instead of dense RPE (one device buffer)
                x.ctrs.record_stream(comp_s)
        mark(staged[slot])

    host_ms = {"upload": 0.0, "forward": 0.0, "download": 0.0}   # host wall time spent enqueueing each phase

    def run_e2e(n_steps):
        upload(0)
        for i in range(n_steps):
            slot = i & 1
            t0 = time.perf_counter()
            if i + 1 < n_steps:
                upload(slot ^ 1)
            t1 = time.perf_counter()
            comp_s.wait_event(ev_in[slot])
            cls, reg, aux = net(staged[slot])                   # the call scenario_tree.py:71 makes
            t2 = time.perf_counter()
            pk = net._last_packed
            if world > 1:                                       # same overlapped collective as in the device-timed loop
                if pending[slot] is not None:
                    pending[slot][0].wait()
                gather_bufs[slot], work = all_gather_packed(pk[6], B, pk[1].shape[0], out=gather_bufs[slot], async_op=True)
                pending[slot] = (work, pk)
            ev_out[slot].record(comp_s)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_out[slot])
                if out_host[slot] is None:
                    out_host[slot] = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in pk[:3]]
                for h, d in zip(out_host[slot], pk[:3]):
                    h.copy_(d, non_blocking=True)               # D2H of cls, reg, vel
                ev_done[slot].record(copy_s)
            for t in pk[:3]:
                t.record_stream(copy_s)
            t3 = time.perf_counter()
            host_ms["upload"] += (t1 - t0) * 1e3
            host_ms["forward"] += (t2 - t1) * 1e3
            host_ms["download"] += (t3 - t2) * 1e3
        drain()
        comp_s.wait_stream(copy_s)
    def measure_e2e(n_steps):
        run_e2e(max(3, args.warmup))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for k in host_ms:
            host_ms[k] = 0.0
        e0.record()
        run_e2e(n_steps)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), {k: round(v / n_steps, 3) for k, v in host_ms.items()}

    # (a) the plug-in as shipped: pre_process reads the anchors out of the collated dict (2.6 KB per scene) and the
    #     device evaluates get_rpe; (b) the same call made to ship the dense RPE tensors like the reference's gpu() does
    net.rpe_on_device = True
    ms2, host_enq = measure_e2e(args.steps)
    e2e = world * B * args.steps / (ms2 * 1e-3)
    common = sum(x.numel() * 4 for x in (host[0], host[2], host[5], host[6]))
    h2d = common + sum(4 * (t["TRAJS_CTRS"].numel() + t["TRAJS_VECS"].numel()) for t in host_full["TRAJS"]) \
        + sum(4 * (g["lane_ctrs"].numel() + g["lane_vecs"].numel()) for g in host_full["LANE_GRAPH"])
    net.rpe_on_device = False
    n_dense = max(3, args.steps // 2)
    ms3, host_enq_dense = measure_e2e(n_dense)
    e2e_dense = {"value": world * B * n_dense / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / n_dense,
                 "h2d_bytes_per_step": common + sum(r["scene"].numel() * 4 for r in host[4]),
                 "host_enqueue_ms_per_step": host_enq_dense,
                 "note": "pre_process made to upload the dense [5,M,M] RPE tensors (rpe_on_device = False)"}
    net.rpe_on_device = True
    out_host = out_host[0]
    d2h = sum(x.numel() * 4 for x in out_host)

    sharded = None
    if world > 1 and not args.no_tree_sharded:
        # tree mode of SURVEY.md 8e on ALL ranks: forced-full depth-4 x branch-6 tree, every level's frontier sharded over the
        # ranks, one all-gather of the packed (cls | reg | vel) buffer per level; wall time, max over ranks
        import copy
        from mind_b200 import synth
        from mind_b200.scenario_tree import ScenarioTreeGeneratorB200
        sargs = dict(x0=(100, 108, 92, 120, 112, 96, 130, 85), y0=(0, 3.5, -3.5, 0, 3.5, 3.5, -3.5, 0), v=(5, 9, 3, 8, 2, 10, 6, 12))
        gen = ScenarioTreeGeneratorB200(dev, net, 50, 50, _TreeCfg())
        gen.force_full, gen.distributed = (10, 20, 30), True
        net.set_option("tc_min_tokens", 0)              # forced decisions: fused tier, as in tree_rollout.forced_full
        times = []
        for r in range(7):
            data, lane, info, graph = synth.scene_s3(**sargs)
            gen.reset(); gen.set_target_lane(lane, info); gen.lane_graph = copy.deepcopy(graph)
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            gen.rollout(data)
            torch.cuda.synchronize()
            t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        net.set_option("tc_min_tokens", 128)
        sharded = {"ms_per_tree": statistics.median(times[2:]), "level_batches": list(gen.net_batches), "ranks": world,
                   "single_gpu_ms_per_tree": None}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    f_ms, f_n = prof.get("fusion_tc", (0.0, 0))
    roof = None
    if f_n:
        per_launch_s = f_ms / f_n * 1e-3
        flops = B * fusion_flops_per_scene(True)
        ach = flops / per_launch_s / 1e12
        hb = B * fusion_bytes_per_scene(True) / per_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "fusion_tc_traffic.json")
        if os.path.exists(tp) and B == 256:                     # dram bytes per launch from the committed ncu --set full capture
            traffic = json.load(open(tp))["dram_bytes_per_launch"]
        roof = {"kernel": "k_rela_fusion_tc (layers 0-4)", "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_sust"], "frac_of_burst_peak": ach / pk["tf_burst"], "peak_burst": pk["tf_burst"],
                "traffic": traffic, "traffic_source": "ncu --set full capture of this kernel build (profiles/fusion_tc_traffic.json)" if traffic else None,
                "peak_source": pk["src"] + ", sustained bf16/fp16 dense (the kernel is timed inside a long step)",
                "ms_per_launch": f_ms / f_n, "launches_timed": f_n,
                "hbm_algorithmic_gbs": hb, "hbm_frac_of_measured": hb / pk["hbm"],
                "flops_per_launch": flops, "bytes_per_launch": B * fusion_bytes_per_scene(True)}
    cpu_v, cores, sample = cpu_port_rate(sd)
    tree = bench_tree(net, dev)
    if sharded is not None:
        sharded["single_gpu_ms_per_tree"] = tree.get("forced_full", {}).get("ms_per_tree")
    cost = bench_cost_fields(dev) if rank == 0 else None
    closed = bench_closed_loop(net, dev) if rank == 0 else None
    eager = gpu_eager_rate(sd, dev) if rank == 0 else None
    if eager and "scenes_per_s_B64" in eager:
        eager["native_over_eager_device_resident"] = value / world / eager["scenes_per_s_B64"]
    stage_ms = {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items())}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if args.precision == "f16tc" else "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: batch=%d synthetic scenes per GPU, 32 actors x 128 lane polylines, K=6" % B,
                       "global_batch": world * B, "precision": args.precision,
                       "l2": "inputs larger than L2 (fp16 edge stream %.2f GB per step)" % (B * N_TOK * N_TOK * 256 / 1e9),
                       "collective": "ONE all_gather per step of the packed cls|reg|vel buffer the decoder wrote, overlapped with the next step" if world > 1 else "none",
                       "weights": "reference checkpoint 20240121-172745"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms2 / args.steps, "host_enqueue_ms_per_step": host_enq,
                    "inputs": "pinned host collated dict -> pre_process (actors, lanes, target, anchors; get_rpe on the device) -> "
                              "net() -> cls/reg/vel in pinned host memory"},
            "e2e_dense_rpe": e2e_dense,
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "stage_ms_per_step": stage_ms,
            "gpu_eager_baseline": eager, "cost_fields": cost, "closed_loop": closed, "tree_rollout_sharded": sharded,
            "tree_rollout": {"unit": "ms/scene", "scene": "natural / forced_full: S3 kinematic, 8 actors x 60 lane polylines; demo_2_*: the Argoverse-2 demo_2 scene of BASELINE.json configs[2]", **tree}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default="f16tc", choices=["f16tc", "fp32"])
    ap.add_argument("--kernel-only", action="store_true", help="device-resident leg only (for ncu runs)")
    ap.add_argument("--no-tree-sharded", action="store_true",
                    help="N > 1: skip the forced-full scenario tree with every level's frontier sharded over the ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
