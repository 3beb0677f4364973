"""AIME scenario-tree generator on the B200 path.

Same constructor and methods as the reference's ScenarioTreeGenerator
(planners/mind/scenario_tree.py:19-58: __init__(device, network, obs_len, pred_len, config),
reset, set_target_lane, branch_aime, get_scenario_tree) and the same host control flow, including
its quirks (depth counter in node ids :107/:321, re-examination of leaves :84-100).  Between two
batched network calls everything numeric runs in libmind_b200.so: one mind_tree_level launch set
per depth level (mode sort, frame transforms, pruning, topology merge, branch-time scan) and one
mind_tree_update (observation re-normalisation, actor features, lane anchors, high-level command);
the host only reads back four small [F,6] decision arrays per level and keeps the tree bookkeeping.
The finished trees are handed out exactly as the reference does: a list of Tree objects whose
node.data = [prob, trajs (Na,dur,2) f32, covs (Na,dur,1) f32, tgt_pts (11,2)] (numpy).
"""
import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import lib as _lib
from . import plumbing as P

try:    # inside a MIND checkout: hand out the reference's own container types
    from planners.basic.tree import Node, Tree          # type: ignore
except Exception:                                       # stand-alone: structural mirror (planners/basic/tree.py)
    class Node:
        def __init__(self, key, parent_key, data):
            self.key, self.parent_key, self.data = key, parent_key, data
            self.children_keys, self.depth = [], 0

        def __str__(self):
            return "Node_%s: Parent: %s, Children: %s" % (self.key, self.parent_key, self.children_keys)

    class Tree:
        """Same interface and leaf ordering as planners/basic/tree.py; leaves are kept in an insertion-ordered
        dict so that attaching a child is O(1) (the reference's list.remove makes big trees quadratic)."""
        def __init__(self):
            self.nodes, self.root, self._leaves = {}, None, {}

        @property
        def leaves(self):
            return list(self._leaves)

        def get_node(self, key):
            return self.nodes[key]

        def get_root(self):
            return self.nodes[self.root]

        def get_root_key(self):
            return self.root

        def has_children(self, key):
            return len(self.nodes[key].children_keys) > 0

        def get_children_keys(self, key):
            return self.nodes[key].children_keys

        def add_node(self, node):
            if node.parent_key is None and not self.nodes:
                self.nodes[node.key] = node
                self.root = node.key
                self._leaves[node.key] = None
                return
            if node.parent_key not in self.nodes:
                raise KeyError("Parent does not exist.")
            if node.key in self.nodes:
                raise ValueError("Node key already exists.")
            self.nodes[node.parent_key].children_keys.append(node.key)
            self._leaves.pop(node.parent_key, None)
            node.depth = self.nodes[node.parent_key].depth + 1
            self.nodes[node.key] = node
            self._leaves[node.key] = None

        def get_leaf_nodes(self):
            return [self.nodes[k] for k in self._leaves]

        def get_leaf_keys(self):
            return list(self._leaves)

        def retrieve_nodes_to_root(self, key):
            out, cur = [], self.nodes[key]
            out.append(cur)
            while cur.parent_key is not None:
                cur = self.nodes[cur.parent_key]
                out.append(cur)
            return out

        def size(self):
            return len(self.nodes)


class _Scen:
    """scenario_tree.py:10-16 with device-side payload replaced by (level, row) handles."""
    def __init__(self, rec, branch_flag=False, end_flag=False, terminate_flag=False):
        self.rec = rec                  # dict: level, row, f, prob, cur_t, end_t, tb0, examined
        self.obs_slot = None            # index of this node's scene in the next frontier
        self.branch_flag, self.end_flag, self.terminate_flag = branch_flag, end_flag, terminate_flag


class _Level:
    """Device state of one frontier (F scenes): inputs of the batched network call + parents' histories."""
    pass


class ScenarioTreeGeneratorB200:
    def __init__(self, device, network, obs_len=50, pred_len=50, config=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScenarioTreeGeneratorB200 needs a CUDA device (no CPU fallback)")
        if obs_len != 50 or pred_len != 50:
            # the tree-step kernels lay child histories out as 100-step rows (50 observed + 50 predicted), which is what
            # MINDPlanner passes (planners/mind/planner.py:20-21,53); anything else would read past a row
            raise ValueError("obs_len / pred_len must be 50 / 50 (got %r / %r)" % (obs_len, pred_len))
        self.network, self.obs_len, self.pred_len = network, obs_len, pred_len
        self.seq_len = obs_len + pred_len
        self.config = config
        self.tree = Tree()
        self.lane_graph = None
        self.target_lane = self.target_lane_info = None
        self.ego_idx = 0
        self.branch_depth = 0
        self.net_batches: List[int] = []
        self._lib = _lib.load()
        self._levels: List[_Level] = []
        self.timing = {}                # seconds per phase of the last rollout (host wall clock)
        self.front_end = None           # callable (lcl_smp, agent_obs) -> collated scene dict; default ArgoFrontEnd
        # benchmark mode: keep all 6 modes of every scene, branch at fixed times
        self.force_full = None          # e.g. (10, 20, 30): children of level d branch at force_full[d]
        # device buffers that feed the network, kept across rollouts per (level, frontier size): stable pointers let
        # the library replay one captured CUDA graph per level instead of ~170 launches (mind_set_option "graph")
        self._pool = {}
        self.graphs = True
        # one process per GPU: shard every level's frontier over the ranks of `group` (torch.distributed, NCCL)
        self.distributed, self.group = False, None
        self._graphs_on = False

    # ---- reference surface ------------------------------------------------------------------
    def reset(self):
        self.branch_depth = 0
        self.tree = Tree()
        self._levels = []
        self.net_batches = []

    def set_target_lane(self, target_lane, target_lane_info):                     # :110-120
        self.target_lane = torch.from_numpy(np.array(target_lane)).float().to(self.device).contiguous()
        self.target_lane_info = P.pack_target_lane_info(target_lane_info).float().to(self.device).contiguous()

    def branch_aime(self, lcl_smp, agent_obs):                                    # :38-58
        data = self.process_data(lcl_smp, agent_obs)
        return self.rollout(data)

    def process_data(self, lcl_smp, agent_obs):
        """:122-206: observation tracks + static map -> collated scene dict.  Default: mind_b200.front_end.ArgoFrontEnd
        (host restatement with the map-dependent work cached per map); any callable (lcl_smp, agent_obs) -> dict can
        be plugged in through `front_end`, e.g. the reference's own bound process_data."""
        if self.front_end is None:
            from .front_end import ArgoFrontEnd
            self.front_end = ArgoFrontEnd(self)
        return self.front_end(lcl_smp, agent_obs)

    def _buf(self, key, name, shape, dtype=torch.float32):
        d = self._pool.setdefault(key, {})
        t = d.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = d[name] = torch.empty(shape, dtype=dtype, device=self.device)
        return t

    def _tick(self, name, t0):
        import time
        self.timing[name] = self.timing.get(name, 0.0) + (time.perf_counter() - t0)
        return time.perf_counter()

    def rollout(self, data):
        """AIME iteration from a collated one-scene dict (what process_data returns).  The cyclic garbage collector is held
        off for the duration of the call: a rollout allocates thousands of small container objects (none of them cyclic),
        and a generation-2 collection triggered in the middle of one costs ~100 ms on a process that has torch loaded."""
        import gc
        was_on = gc.isenabled()
        gc.disable()
        try:
            return self._rollout(data)
        finally:
            if was_on:
                gc.enable()

    def _rollout(self, data):
        import time
        self.timing = {}
        t = time.perf_counter()
        self.init_scenario_tree(data)
        t = self._tick("root_level", t)
        nodes = self.get_branch_set()
        guard = 0
        while nodes:
            level = self._levels[-1]
            self.predict_scenes(level)
            t = self._tick("predict", t)
            pm = self.prune_merge(level, nodes)
            t = self._tick("tree_level+d2h", t)
            self.create_nodes(pm)
            t = self._tick("create_nodes", t)
            self.decide_branch()
            t = self._tick("decide+update", t)
            nodes = self.get_branch_set()
            guard += 1
            if guard > 64:
                raise RuntimeError("scenario tree did not converge (re-expansion loop, reference quirk :84-100)")
        assert len(self.get_end_set()) > 0, "No end node found in the scenario tree."
        out = self.get_scenario_tree()
        self._tick("pack_output", t)
        return out

    # ---- level construction -----------------------------------------------------------------
    def _root_level(self, data, L: "_Level") -> "_Level":
        """prepare_root_data (:414-465): observations (actor-local) -> global-frame histories.  One scene of a few
        actors: the frame arithmetic runs on the host (same torch fp32 ops as the reference) and everything the device
        needs travels in ONE pinned buffer / one H2D copy instead of ~30 small copies and ~25 small launches.  Called
        AFTER the root scene's network forward has been enqueued (the forward needs none of this), so the ~1 ms of host
        arithmetic runs under the GPU's work instead of in front of it."""
        dev = self.device
        cpu = lambda t: t.detach().to("cpu", torch.float32)
        tj = data["TRAJS"][0]
        orig, rot = cpu(data["ORIG"][0]), cpu(data["ROT"][0])
        ctrs, vecs = cpu(tj["TRAJS_CTRS"]), cpu(tj["TRAJS_VECS"])
        th_g = torch.atan2(rot[1, 0], rot[0, 0])
        th = torch.atan2(vecs[:, 1], vecs[:, 0])
        R = torch.stack([torch.cos(th), -torch.sin(th), torch.sin(th), torch.cos(th)], 1).view(-1, 2, 2)
        pos = torch.matmul(cpu(tj["TRAJS_POS_OBS"]), R.transpose(-1, -2)) + ctrs[:, None]
        vel = torch.matmul(cpu(tj["TRAJS_VEL_OBS"]), R.transpose(-1, -2))
        ang_obs = cpu(tj["TRAJS_ANG_OBS"])
        ang = torch.atan2(ang_obs[..., 1], ang_obs[..., 0])
        Na = pos.shape[0]
        if Na > 256:
            raise ValueError("the tree-step kernels support at most 256 actors per scene (got %d)" % Na)
        if pos.shape[1] != self.obs_len:
            raise ValueError("observation length %d != obs_len %d" % (pos.shape[1], self.obs_len))
        graph = self.lane_graph if self.lane_graph is not None else data["LANE_GRAPH"][0]
        parts = [("hpos", (torch.matmul(pos, rot.T) + orig), (1, Na, 50, 2)),
                 ("hvel", torch.matmul(vel, rot.T), (1, Na, 50, 2)),
                 ("hang", ang + th[:, None] + th_g, (1, Na, 50)),
                 ("hcov", torch.full((1, Na, 50), 1e-5), (1, Na, 50)),
                 ("orig", orig, (1, 2)), ("rot", rot, (1, 4)), ("ctrs", ctrs, (1, Na, 2)), ("vecs", vecs, (1, Na, 2)),
                 ("pprob", torch.ones(1), (1,)),
                 ("tgt_pts", cpu(data["TGT_PTS"][0]), (1, 11, 2)),
                 ("ttype", cpu(tj["TRAJS_TYPE"]), None),           # [Na,50,7]: carried per step into every child scene (:486,524)
                 ("lane_ctrs", cpu(graph["lane_ctrs"]), None), ("lane_vecs", cpu(graph["lane_vecs"]), None)]
        flat = [t.reshape(-1) for _, t, _ in parts]
        sizes = [((f.numel() + 63) // 64) * 64 for f in flat]            # 256-byte aligned pieces
        total = sum(sizes)
        key = ("rootpack", Na, int(flat[-1].numel()))
        pool = self._pool.setdefault(key, {})
        if "pin" not in pool or pool["pin"].numel() < total:
            pool["pin"] = torch.empty(total, dtype=torch.float32).pin_memory()
            pool["dev"] = torch.empty(total, dtype=torch.float32, device=dev)
        pin, dbuf = pool["pin"], pool["dev"]
        off, view = 0, {}
        for (name, t, shape), f, n in zip(parts, flat, sizes):
            pin[off:off + f.numel()].copy_(f)
            view[name] = dbuf[off:off + f.numel()].view(shape if shape is not None else tuple(t.shape))
            off += n
        dbuf[:total].copy_(pin[:total], non_blocking=True)
        L.hpos, L.hvel, L.hang, L.hcov = view["hpos"], view["hvel"], view["hang"], view["hcov"]
        L.orig, L.rot, L.ctrs, L.vecs = view["orig"], view["rot"], view["ctrs"], view["vecs"]
        L.pprob = view["pprob"]
        L.cur_t = self._buf(key, "cur_t0", (1,), torch.int32)
        if "cur_t0_set" not in pool:
            L.cur_t.zero_()
            pool["cur_t0_set"] = True
        L.cur_t_host = [0]
        L.tgt_pts = view["tgt_pts"]
        L.parent_keys = ["root"]
        # constants of the tree
        self._ttype = view["ttype"]
        self._lane_ctrs, self._lane_vecs = view["lane_ctrs"], view["lane_vecs"]
        return L

    def _root_inputs(self, data, d, Na):
        """network.pre_process for the root scene (:69), into persistent buffers (one scene: actors, lanes, RPE, target)."""
        if self.graphs and not self._graphs_on and hasattr(self.network, "use_graphs"):
            self.network.use_graphs(True)
            self._graphs_on = True
        rpe = data["RPE"][0]
        rpe = rpe["scene"] if isinstance(rpe, dict) else rpe
        src = (("actors", data["ACTORS"]), ("lanes", data["LANES"]), ("rpe", rpe), ("tgt_nodes", data["TGT_NODES"]),
               ("tgt_rpe", data["TGT_RPE"]))
        key = ("root", Na, int(data["LANES"].shape[0]))
        out = {}
        for name, t in src:
            b = self._buf(key, name, t.shape)
            b.copy_(t, non_blocking=True)
            out[name] = b
        return (out["actors"], data["ACTOR_IDCS"], out["lanes"], data["LANE_IDCS"], [{"scene": out["rpe"], "scene_mask": None}],
                out["tgt_nodes"], out["tgt_rpe"])

    def init_scenario_tree(self, data):                                           # :60-67
        Na = int(data["TRAJS"][0]["TRAJS_POS_OBS"].shape[0])
        root = _Level()
        root.F, root.Na, root.geom = 1, Na, None
        root.net_in = self._root_inputs(data, None, Na)
        self._lanes = root.net_in[2]
        self._n_lane = self._lanes.shape[0]
        self._levels.append(root)
        rn = Node("root", None, _Scen(None, branch_flag=True))
        rn.data.obs_slot = 0
        self.tree.add_node(rn)
        self.predict_scenes(root)                 # the GPU works on the root prediction ...
        self._root_level(data, root)              # ... while the host builds the global-frame histories (one packed upload)
        self.create_nodes(self.prune_merge(root, [rn]))
        self.decide_branch()

    def predict_scenes(self, level: _Level):                                      # :69-71 (one batched call per level)
        self.net_batches.append(level.F)
        if self.distributed:
            # tree mode of SURVEY.md 8e: frontier sharded over ranks, one all-gather of (cls, reg, vel) per level;
            # frontiers smaller than the world are predicted replicated
            from .distributed import sharded_level_forward
            fwd = lambda net_in, geom: self.network.forward_packed(net_in, geom=geom, persistent_out=True)
            pk = sharded_level_forward(fwd, level.net_in, level.geom, level.F, self.group)
        else:
            pk = self.network.forward_packed(level.net_in, geom=level.geom, persistent_out=True)
        level.cls, level.reg, level.vel = pk[0], pk[1], pk[2]
        return pk

    # ---- prune & merge (:281-412) + branch-time scan (:592-611), whole level on the device -----
    def prune_merge(self, level: _Level, nodes):
        dev, F, Na = self.device, level.F, level.Na
        f32 = dict(device=dev, dtype=torch.float32)
        lv_index = next(i for i, L in enumerate(self._levels) if L is level)
        ckey = ("children", lv_index, F, Na)                            # child histories: device buffers kept across rollouts
        level.cpos = self._buf(ckey, "cpos", (F, 6, Na, 100, 2))
        level.cvel = self._buf(ckey, "cvel", (F, 6, Na, 100, 2))
        level.cang = self._buf(ckey, "cang", (F, 6, Na, 100))
        level.ccov = self._buf(ckey, "ccov", (F, 6, Na, 100))
        level.gpos = self._buf(ckey, "gpos", (F, 6, Na, 60, 2))
        ibuf = self._buf(ckey, "ibuf", (4, F, 6), torch.int32)          # order, keep, tb | cprob (fp32 bits): one D2H
        cprob = ibuf[3].view(torch.float32)
        a = _lib.MindTreeLevel()
        a.n_frontier, a.n_actor, a.obs_len, a.pred_len = F, Na, self.obs_len, self.pred_len
        a.ego_idx = self.ego_idx if self.ego_idx is not None else -1
        a.n_tlane = 0 if self.target_lane is None else self.target_lane.shape[0]
        a.tar_dist_thres = float(self.config.tar_dist_thres)
        for name, t in (("cls", level.cls), ("reg", level.reg), ("vel", level.vel), ("orig", level.orig), ("rot", level.rot),
                        ("ctrs", level.ctrs), ("vecs", level.vecs), ("hpos", level.hpos), ("hang", level.hang),
                        ("hvel", level.hvel), ("hcov", level.hcov), ("pprob", level.pprob), ("cur_t", level.cur_t),
                        ("tlane", self.target_lane), ("cpos", level.cpos), ("cang", level.cang), ("cvel", level.cvel),
                        ("ccov", level.ccov), ("gpos", level.gpos), ("order", ibuf[0]), ("cprob", cprob), ("keep", ibuf[1]),
                        ("tb", ibuf[2])):
            setattr(a, name, t.data_ptr() if t is not None else None)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if self._lib.mind_tree_level(C.byref(a), C.c_void_p(stream)) != 0:
            raise RuntimeError(self._lib.mind_tree_last_error().decode())
        ih = ibuf.cpu().numpy()                                           # the level's only D2H: decisions
        ph = ih[3].view(np.float32)
        if self.force_full is not None:
            ih[1] = 1
            ih[2] = self.force_full[lv_index] if lv_index < len(self.force_full) else self.pred_len
        # kept children in (scene, slot) order; the bookkeeping below is plain Python on ~F*6 small records
        fs, ks = np.nonzero(ih[1])
        modes, tbs, probs = ih[0][fs, ks].tolist(), ih[2][fs, ks].tolist(), ph[fs, ks]
        depth, end_t, cur = self.branch_depth, self.pred_len, level.cur_t_host
        keys = [n.key for n in nodes]
        return [dict(SCEN_ID="%d_%d_%d" % (depth, f, m), PARENT_ID=keys[f], level=lv_index, row=f * 6 + k, f=f, prob=p,
                     cur_t=cur[f], end_t=end_t, tb0=tb, examined=False)
                for f, k, m, tb, p in zip(fs.tolist(), ks.tolist(), modes, tbs, probs)]

    def create_nodes(self, preds):                                                # :73-80
        for p in preds:
            self.tree.add_node(Node(p["SCEN_ID"], p["PARENT_ID"], _Scen(p)))

    def get_branch_time(self, rec):                                               # :592-611
        if not rec["examined"]:
            rec["examined"] = True
            if rec["tb0"] < rec["end_t"]:
                rec["end_t"] = rec["tb0"]
        return rec["end_t"]

    def decide_branch(self):                                                      # :82-100
        to_branch = []
        for l in self.tree.get_leaf_nodes():
            s = l.data
            if s.branch_flag:
                s.branch_flag = False
                s.terminate_flag = True
            elif not s.end_flag:
                if l.depth >= self.config.max_depth:
                    s.terminate_flag = True
                else:
                    t_b = self.get_branch_time(s.rec)
                    if t_b < self.pred_len:
                        to_branch.append(l)
                        s.branch_flag = True
                    else:
                        s.end_flag = True
        if to_branch:
            self._levels.append(self.update_obser(to_branch))

    def get_branch_set(self):                                                     # :102-108
        out = [l for l in self.tree.get_leaf_nodes() if l.data.branch_flag]
        self.branch_depth += 1
        return out

    def get_end_set(self):                                                        # :274-279
        return [n for n in self.tree.get_leaf_nodes() if n.data.end_flag]

    # ---- observation update for every branching child of the level (:467-567) ------------------
    def update_obser(self, nodes) -> _Level:
        dev, G = self.device, len(nodes)
        Na, Nl = self._levels[0].Na, self._n_lane
        f32 = dict(device=dev, dtype=torch.float32)
        N = _Level()
        N.F, N.Na = G, Na
        key = ("level", len(self._levels), G, Na, Nl)
        B_ = lambda name, *shape: self._buf(key, name, shape)
        N.hpos, N.hvel = B_("hpos", G, Na, 50, 2), B_("hvel", G, Na, 50, 2)
        N.hang, N.hcov = B_("hang", G, Na, 50), B_("hcov", G, Na, 50)
        N.orig, N.rot = B_("orig", G, 2), B_("rot", G, 4)
        N.ctrs, N.vecs = B_("ctrs", G, Na, 2), B_("vecs", G, Na, 2)
        actors = B_("actors", G * Na, 14, 48)
        geom_c, geom_v = B_("geom_c", G, Na + Nl, 2), B_("geom_v", G, Na + Nl, 2)
        tgt_nodes, tgt_rpe = B_("tgt_nodes", G, 10, 16), B_("tgt_rpe", G, 20)
        N.tgt_pts = B_("tgt_pts", G, 11, 2)
        N.pprob = torch.tensor([float(n.data.rec["prob"]) for n in nodes], **f32)
        N.cur_t_host = [int(n.data.rec["end_t"]) for n in nodes]
        N.cur_t = torch.tensor(N.cur_t_host, dtype=torch.int32, device=dev)
        N.parent_keys = [n.key for n in nodes]
        stream = torch.cuda.current_stream(dev).cuda_stream
        by_level = {}
        for g, n in enumerate(nodes):
            by_level.setdefault(n.data.rec["level"], []).append(g)
        for lv, gs in by_level.items():
            if gs != list(range(gs[0], gs[0] + len(gs))):
                raise RuntimeError("frontier nodes of one source level must be contiguous")
            src_level = self._levels[lv]
            src = torch.tensor([[nodes[g].data.rec["row"], nodes[g].data.rec["end_t"] - nodes[g].data.rec["cur_t"]] for g in gs],
                               dtype=torch.int32, device=dev)
            g0, n_new = gs[0], len(gs)
            u = _lib.MindTreeUpdate()
            u.n_new, u.n_actor, u.n_lane, u.n_tlane = n_new, Na, Nl, self.target_lane.shape[0]
            u.tar_time_ahead = float(self.config.tar_time_ahead)
            for name, t in (("src", src), ("cpos", src_level.cpos), ("cang", src_level.cang), ("cvel", src_level.cvel),
                            ("ccov", src_level.ccov), ("ttype", self._ttype), ("lane_ctrs", self._lane_ctrs),
                            ("lane_vecs", self._lane_vecs), ("tlane", self.target_lane), ("tinfo", self.target_lane_info),
                            ("npos", N.hpos[g0:]), ("nang", N.hang[g0:]), ("nvel", N.hvel[g0:]), ("ncov", N.hcov[g0:]),
                            ("norig", N.orig[g0:]), ("nrot", N.rot[g0:]), ("nctrs", N.ctrs[g0:]), ("nvecs", N.vecs[g0:]),
                            ("actors", actors[g0 * Na:]), ("geom_c", geom_c[g0:]), ("geom_v", geom_v[g0:]),
                            ("tgt_nodes", tgt_nodes[g0:]), ("tgt_rpe", tgt_rpe[g0:]), ("tgt_pts", N.tgt_pts[g0:])):
                setattr(u, name, t.data_ptr())
            if self._lib.mind_tree_update(C.byref(u), C.c_void_p(stream)) != 0:
                raise RuntimeError(self._lib.mind_tree_last_error().decode())
            self._keep = src
        idc = self._pool[key].get("idcs")
        if idc is None:                                       # only their lengths matter to the network
            idc = self._pool[key]["idcs"] = ([range(g * Na, (g + 1) * Na) for g in range(G)],
                                             [range(g * Nl, (g + 1) * Nl) for g in range(G)])
        a_idcs, l_idcs = idc
        lanes = B_("lanes", G * Nl, 10, 16)
        lanes.view(G, Nl, 10, 16).copy_(self._lanes.unsqueeze(0).expand(G, -1, -1, -1))
        N.net_in = (actors, a_idcs, lanes, l_idcs, None, tgt_nodes, tgt_rpe)
        N.geom = (geom_c.view(-1, 2), geom_v.view(-1, 2))
        for g, n in enumerate(nodes):
            n.data.obs_slot = g
        return N

    # ---- output packing (:208-272) -------------------------------------------------------------
    def get_scenario_tree(self):
        """One pass over the finished tree: label the branches that reached an end node, renormalise sibling
        probabilities, and build the reference's output directly -- one Tree per labelled child of the root with
        node.data = [prob, trajs (Na,dur,2), covs (Na,dur,1), tgt_pts (11,2)] -- while the payloads of all contributing
        levels travel in one asynchronous D2H per level into pinned buffers.  The numpy payloads are views of those
        buffers; two sets alternate, so a result stays valid until the rollout after the next one."""
        tree = self.tree
        get = tree.nodes.__getitem__
        root = tree.get_root()
        for n in self.get_end_set():                                   # label the branches that finished
            n = get(n.parent_key) if n.parent_key is not None else n
            while n.parent_key is not None and not n.data.end_flag:   # ancestors already labelled: stop early
                n.data.end_flag = True
                n = get(n.parent_key)
        trees, todo, levels = [], [], set()
        for key in root.children_keys:
            n = get(key)
            if not n.data.end_flag:
                continue
            st = Tree()
            dn = Node(n.key, None, [1.0])
            st.add_node(dn)
            todo.append((dn, n.data.rec))
            levels.add(n.data.rec["level"])
            queue = [(n, 1.0)]
            while queue:
                c, pp = queue.pop(0)
                kids = [k for k in map(get, c.children_keys) if k.data.end_flag]
                total = 0.0
                for k in kids:
                    total += np.asarray(k.data.rec["prob"])
                for k in kids:
                    pk = np.asarray(k.data.rec["prob"]) / total * pp
                    dn = Node(k.key, c.key, [pk])
                    st.add_node(dn)
                    todo.append((dn, k.data.rec))
                    levels.add(k.data.rec["level"])
                    queue.append((k, pk))
            trees.append(st)
        # one asynchronous D2H per contributing level (predicted half of the child histories only), a single synchronisation
        self._host_flip = getattr(self, "_host_flip", 0) ^ 1
        host = {}
        for lv in sorted(levels):
            L = self._levels[lv]
            ent = []
            for name, t in (("cpos", L.cpos[:, :, :, self.obs_len:]), ("ccov", L.ccov[:, :, :, self.obs_len:]), ("tgt", L.tgt_pts)):
                hk = ("host", self._host_flip, lv, name)
                h = self._pool.get(hk)
                if h is None or tuple(h.shape) != tuple(t.shape):
                    h = self._pool[hk] = torch.empty(tuple(t.shape), dtype=t.dtype).pin_memory()
                h.copy_(t, non_blocking=True)
                ent.append(h)
            host[lv] = ent
        torch.cuda.current_stream(self.device).synchronize()
        host = {lv: tuple(h.numpy() for h in ent) for lv, ent in host.items()}
        for dn, rec in todo:                                           # every labelled node gets its payload once
            dur = rec["end_t"] - rec["cur_t"]
            cpos, ccov, tgt = host[rec["level"]]
            f, k = divmod(rec["row"], 6)
            dn.data += [cpos[f, k, :, :dur, :], ccov[f, k, :, :dur, None], tgt[f]]
        return trees
