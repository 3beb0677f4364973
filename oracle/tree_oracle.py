"""CPU restatement of MIND's AIME scenario-tree generator.  TEST INFRASTRUCTURE ONLY.

Restates planners/mind/scenario_tree.py (ScenarioTreeGenerator) + planners/basic/tree.py on plain
CPU torch, with any network object exposing pre_process / __call__ (the oracle predictor or the
reference's own ScenePredNet).  process_data (av2 / shapely front end, scenario_tree.py:122-206)
is out of scope: the tree starts from the collated scene dict.

Pinned against the UNMODIFIED reference class in tests/test_oracle_vs_reference.py and against
tests/golden/tree_s3.npz (oracle/make_golden_tree.py).
"""
import copy
import math

import torch

from mind_b200 import plumbing as P


class Node:                      # planners/basic/tree.py:1-11
    def __init__(self, key, parent_key, data):
        self.key, self.parent_key, self.data = key, parent_key, data
        self.children_keys, self.depth = [], 0


class Tree:                      # planners/basic/tree.py:14-110 (the subset the path uses)
    def __init__(self):
        self.nodes, self.root, self.leaves = {}, None, []

    def add_node(self, node):
        if node.parent_key is None and not self.nodes:
            self.nodes[node.key] = node
            self.root = node.key
            self.leaves.append(node.key)
            return
        if node.parent_key not in self.nodes:
            raise KeyError("Parent does not exist.")
        if node.key in self.nodes:
            raise ValueError("Node key already exists.")
        self.nodes[node.parent_key].children_keys.append(node.key)
        if node.parent_key in self.leaves:
            self.leaves.remove(node.parent_key)
        node.depth = self.nodes[node.parent_key].depth + 1
        self.nodes[node.key] = node
        self.leaves.append(node.key)

    def get_node(self, key):
        return self.nodes[key]

    def get_root(self):
        return self.nodes[self.root]

    def get_leaf_nodes(self):
        return [self.nodes[k] for k in self.leaves]

    def size(self):
        return len(self.nodes)


class Scen:                      # scenario_tree.py:10-16
    def __init__(self, data, obs_data, branch_flag=False, end_flag=False, terminate_flag=False):
        self.data, self.obs_data = data, obs_data
        self.branch_flag, self.end_flag, self.terminate_flag = branch_flag, end_flag, terminate_flag


def wrap(a):
    return torch.atan2(torch.sin(a), torch.cos(a))


def dist_to_polyline(poly, pt):
    """utils.py:486-513: min over segments of the distance to the clamped projection."""
    p1, p2 = poly[:-1], poly[1:]
    seg = p2 - p1
    t = torch.clamp(((pt - p1) * seg).sum(-1) / (seg * seg).sum(-1), 0, 1)
    return torch.norm(p1 + t[:, None] * seg - pt, dim=-1).min()


class TreeOracle:
    def __init__(self, network, obs_len=50, pred_len=50, config=None, device="cpu"):
        self.network, self.obs_len, self.pred_len = network, obs_len, pred_len
        self.seq_len = obs_len + pred_len
        self.config = config
        self.device = torch.device(device)
        self.tree = Tree()
        self.lane_graph = None
        self.target_lane = self.target_lane_info = None
        self.ego_idx = 0
        self.branch_depth = 0
        self.net_batches = []

    def reset(self):
        self.branch_depth = 0
        self.merge_margins = []
        self.prune_margins = []
        self.tree = Tree()
        self.net_batches = []

    def set_target_lane(self, target_lane, target_lane_info):            # :110-120
        import numpy as np
        self.target_lane = torch.from_numpy(np.array(target_lane))
        self.target_lane_info = P.pack_target_lane_info(target_lane_info)

    # ---- driver (:38-58, without process_data) ----
    def rollout(self, data):
        self.init_scenario_tree(data)
        nodes = self.get_branch_set()
        while nodes:
            batch = P.collate_scenes([n.data.obs_data for n in nodes])
            pred = self.predict_scenes(batch)
            self.create_nodes(self.prune_merge(batch, pred))
            self.decide_branch()
            nodes = self.get_branch_set()
        assert len(self.get_end_set()) > 0, "No end node found in the scenario tree."
        return self.get_scenario_tree()

    def init_scenario_tree(self, data):                                   # :60-67
        root = self.prepare_root_data(data)
        self.tree.add_node(Node("root", None, Scen(None, root, branch_flag=True)))
        pred = self.predict_scenes(root)
        self.create_nodes(self.prune_merge(root, pred))
        self.decide_branch()

    def predict_scenes(self, data):                                       # :69-71
        self.net_batches.append(len(data["ORIG"]))
        return self.network(self.network.pre_process(data))

    def create_nodes(self, preds):                                        # :73-80
        for p in preds:
            self.tree.add_node(Node(p["SCEN_ID"], p["PARENT_ID"], Scen(p, None)))

    def decide_branch(self):                                              # :82-100
        for l in self.tree.get_leaf_nodes():
            if l.data.branch_flag:
                l.data.branch_flag = False
                l.data.terminate_flag = True
            elif not l.data.end_flag:
                if l.depth >= self.config.max_depth:
                    l.data.terminate_flag = True
                else:
                    t_b = self.get_branch_time(l.data.data)
                    if t_b < self.pred_len:
                        l.data.obs_data, l.data.data = self.update_obser(l.data.data)
                        l.data.branch_flag = True
                    else:
                        l.data.end_flag = True

    def get_branch_set(self):                                             # :102-108 (depth counter quirk)
        out = [l for l in self.tree.get_leaf_nodes() if l.data.branch_flag]
        self.branch_depth += 1
        return out

    def get_end_set(self):                                                # :274-279
        return [n for n in self.tree.get_leaf_nodes() if n.data.end_flag]

    # ---- root data (:414-465) ----
    def prepare_root_data(self, data):
        B = len(data["ORIG"])
        for k in ("TRAJS_POS_HIST", "TRAJS_ANG_HIST", "TRAJS_VEL_HIST", "TRAJS_COV_HIST"):
            data[k] = [None] * B
        data["SCEN_PROB"] = [1.0] * B
        data["SCEN_ID"] = ["root"] * B
        data["PARENT_ID"] = [None] * B
        data["CUR_T"] = [0] * B
        data["END_T"] = [self.pred_len] * B
        for b in range(B):
            orig, rot = data["ORIG"][b], data["ROT"][b]
            tj = data["TRAJS"][b]
            ctrs, vecs = tj["TRAJS_CTRS"], tj["TRAJS_VECS"]
            th_g = torch.atan2(rot[1, 0], rot[0, 0])
            th = torch.atan2(vecs[:, 1], vecs[:, 0])
            R = torch.stack([torch.cos(th), -torch.sin(th), torch.sin(th), torch.cos(th)], 1).view(-1, 2, 2)
            pos = torch.matmul(tj["TRAJS_POS_OBS"], R.transpose(-1, -2)) + ctrs[:, None]
            vel = torch.matmul(tj["TRAJS_VEL_OBS"], R.transpose(-1, -2))
            ang_obs = torch.atan2(tj["TRAJS_ANG_OBS"][..., 1], tj["TRAJS_ANG_OBS"][..., 0])
            data["TRAJS_POS_HIST"][b] = torch.matmul(pos, rot.T) + orig
            data["TRAJS_VEL_HIST"][b] = torch.matmul(vel, rot.T)
            data["TRAJS_ANG_HIST"][b] = ang_obs + th[:, None] + th_g
            data["TRAJS_COV_HIST"][b] = 1e-5 * torch.ones(pos.shape[0], pos.shape[1], 1)
        return data

    # ---- prune & merge (:281-412) ----
    def prune_merge(self, data, out):
        res = []
        cls_b, reg_b, aux_b = out
        for b in range(len(data["ORIG"])):
            orig, rot = data["ORIG"][b], data["ROT"][b]
            tj = data["TRAJS"][b]
            ctrs, vecs = tj["TRAJS_CTRS"], tj["TRAJS_VECS"]
            th_g = torch.atan2(rot[1, 0], rot[0, 0])
            reg = reg_b[b].detach().cpu().clone()
            cls = cls_b[b].detach().cpu()
            vel_all = aux_b[b][0].detach().cpu().clone()
            ang_all = torch.atan2(vel_all[..., 1], vel_all[..., 0])        # res_ang from the un-rotated vel (:311)
            order = torch.argsort(cls, dim=1, descending=True)[0]
            th = torch.atan2(vecs[:, 1], vecs[:, 0])
            R = torch.stack([torch.cos(th), -torch.sin(th), torch.sin(th), torch.cos(th)], 1).view(-1, 2, 2)
            cands = []
            for m in order:
                m = int(m)
                prob = cls[0, m]
                pos = torch.matmul(reg[:, m, :, :2], R.transpose(-1, -2)) + ctrs[:, None]
                vel = torch.matmul(vel_all[:, m], R.transpose(-1, -2))
                cov = torch.maximum(reg[:, m, :, 2], reg[:, m, :, 3]).unsqueeze(-1)
                pos = torch.matmul(pos, rot.T) + orig
                vel = torch.matmul(vel, rot.T)
                ang = ang_all[:, m] + th[:, None] + th_g
                cov = cov + data["TRAJS_COV_HIST"][b][:, -1].unsqueeze(1)
                cur = dict(SCEN_PROB=prob * data["SCEN_PROB"][b], CUR_T=data["CUR_T"][b], END_T=data["END_T"][b],
                           PARENT_ID=data["SCEN_ID"][b], SCEN_ID="{}_{}_{}".format(self.branch_depth, b, m),
                           TRAJS=dict(TRAJS_TYPE=tj["TRAJS_TYPE"], TRAJS_TID=tj["TRAJS_TID"], TRAJS_CAT=tj["TRAJS_CAT"]),
                           TRAJS_POS_HIST=torch.cat([data["TRAJS_POS_HIST"][b], pos], 1)[:, :self.seq_len],
                           TRAJS_COV_HIST=torch.cat([data["TRAJS_COV_HIST"][b], cov], 1)[:, :self.seq_len],
                           TRAJS_ANG_HIST=torch.cat([data["TRAJS_ANG_HIST"][b], ang], 1)[:, :self.seq_len],
                           TRAJS_VEL_HIST=torch.cat([data["TRAJS_VEL_HIST"][b], vel], 1)[:, :self.seq_len],
                           TGT_PTS=data["TGT_PTS"][b])
                # test infrastructure: distance of every prune decision from its threshold (relative for the probability,
                # metres for the target-lane test), next to merge_margins
                self.prune_margins.append((self.branch_depth, b, "prob", float(cur["SCEN_PROB"]) / 0.001 - 1.0))
                if cur["SCEN_PROB"] < 0.001:
                    continue
                if self.target_lane is not None and self.ego_idx is not None:
                    ego_mean = cur["TRAJS_POS_HIST"][self.ego_idx][-1]
                    ego_cov = cur["TRAJS_COV_HIST"][self.ego_idx][-1]
                    over = dist_to_polyline(self.target_lane, ego_mean) - ego_cov - self.config.tar_dist_thres
                    self.prune_margins.append((self.branch_depth, b, "lane", float(over)))
                    if dist_to_polyline(self.target_lane, ego_mean) - ego_cov > self.config.tar_dist_thres:
                        continue
                rel = pos[1:] - pos[0:1]                                   # exo - ego over the 60 predicted steps
                rel = rel / torch.norm(rel, dim=-1, keepdim=True)
                a = torch.atan2(rel[..., 1], rel[..., 0])
                topo = wrap(a[:, 1:] - a[:, :-1]).sum(dim=1)
                cands.append((cur, topo))
            while cands:                                                   # greedy merge (:396-410)
                sel, st = cands[0]
                res.append(sel)
                # test infrastructure: how far each keep / merge decision sits from the pi/6 threshold (rad); a tree whose
                # smallest |margin| is below the noise of the inputs has no well-defined node set
                self.merge_margins += [(self.branch_depth, b, float((wrap(st - c[1]).abs() - math.pi / 6).max())) for c in cands[1:]]
                cands = [c for c in cands[1:] if bool(((wrap(st - c[1]).abs() - math.pi / 6) > 0).sum() > 0)]
        return res

    # ---- branching decision (:592-611) ----
    def get_branch_time(self, d):
        cov, cur_t, end_t = d["TRAJS_COV_HIST"], d["CUR_T"], d["END_T"]
        cmp_t = self.obs_len + cur_t + (1 if cur_t == 0 else 0)
        for t in range(cur_t + 1, end_t):
            if t % 2 == 1:
                continue
            if bool((cov[:, self.obs_len + t] / cov[:, cmp_t] > 9).sum() > 0):
                d["END_T"] = t
                return t
        return end_t

    # ---- observation update (:467-567) ----
    def update_obser(self, cur):
        dur = cur["END_T"] - cur["CUR_T"]
        for k in ("TRAJS_POS_HIST", "TRAJS_COV_HIST", "TRAJS_ANG_HIST", "TRAJS_VEL_HIST"):
            cur[k] = cur[k][:, :self.obs_len + dur]
        d = copy.deepcopy(cur)
        d["CUR_T"], d["END_T"] = cur["END_T"], self.pred_len
        for k in ("TRAJS_POS_HIST", "TRAJS_COV_HIST", "TRAJS_ANG_HIST", "TRAJS_VEL_HIST"):
            d[k] = d[k][:, -self.obs_len:]
        pos, ang, vel = d["TRAJS_POS_HIST"], d["TRAJS_ANG_HIST"], d["TRAJS_VEL_HIST"]
        orig, rot, theta = P.origin_rotation(pos[0], ang[0])
        pos = torch.matmul(pos - orig, rot)
        ang = ang - theta
        vel = torch.matmul(vel, rot)
        pn, an, vn, ctrs, vecs = [], [], [], [], []
        for i in range(pos.shape[0]):
            o, r, th = P.origin_rotation(pos[i], ang[i])
            pn.append(torch.matmul(pos[i] - o, r))
            an.append(ang[i] - th)
            vn.append(torch.matmul(vel[i], r))
            ctrs.append(o)
            vecs.append(torch.stack([torch.cos(th), torch.sin(th)]))
        an = torch.stack(an)
        vn = torch.stack(vn)
        trajs = dict(TRAJS_POS_OBS=torch.stack(pn), TRAJS_ANG_OBS=torch.stack([torch.cos(an), torch.sin(an)], -1),
                     TRAJS_VEL_OBS=vn, TRAJS_TYPE=d["TRAJS"]["TRAJS_TYPE"], PAD_OBS=torch.ones_like(an)[:, :self.obs_len],
                     TRAJS_CTRS=torch.stack(ctrs), TRAJS_VECS=torch.stack(vecs), TRAJS_TID=d["TRAJS"]["TRAJS_TID"],
                     TRAJS_CAT=d["TRAJS"]["TRAJS_CAT"])
        g = copy.deepcopy(self.lane_graph)                                 # utils.py:171-177
        g["lane_ctrs"] = torch.matmul(g["lane_ctrs"] - orig, rot)
        g["lane_vecs"] = torch.matmul(g["lane_vecs"], rot)
        rpe = {"scene": P.pairwise_rpe(torch.cat([trajs["TRAJS_CTRS"], g["lane_ctrs"]]),
                                       torch.cat([trajs["TRAJS_VECS"], g["lane_vecs"]])), "scene_mask": None}
        tgt_pts, tgt_nodes, anch = P.high_level_command(self.target_lane, self.target_lane_info, orig, rot,
                                                        vn[0, -1].norm(), self.config.tar_time_ahead)
        tgt_rpe = P.pairwise_rpe(torch.stack([anch[0], trajs["TRAJS_CTRS"][0]]), torch.stack([anch[1], trajs["TRAJS_VECS"][0]]))
        d.update(ORIG=orig, ROT=rot, TRAJS=trajs, LANE_GRAPH=g, RPE=rpe, TGT_PTS=tgt_pts, TGT_NODES=tgt_nodes,
                 TGT_ANCH=anch, TGT_RPE=tgt_rpe)
        return d, cur

    # ---- output packing (:208-272) ----
    def get_scenario_tree(self):
        dt = Tree()
        root = self.tree.get_root()
        dt.add_node(Node(root.key, None, [1.0]))
        for n in self.get_end_set():
            while n.parent_key is not None:
                n.data.end_flag = True
                n = self.tree.get_node(n.parent_key)
        for key in root.children_keys:
            n = self.tree.get_node(key)
            if not n.data.end_flag:
                continue
            dt.add_node(Node(n.key, root.key, [1.0]))
            queue = [n]
            while queue:
                c = queue.pop(0)
                pp = dt.get_node(c.key).data[0]
                kids = [self.tree.get_node(k) for k in c.children_keys if self.tree.get_node(k).data.end_flag]
                total = 0.0
                for k in kids:
                    total += k.data.data["SCEN_PROB"].cpu().numpy()
                for k in kids:
                    dt.add_node(Node(k.key, c.key, [k.data.data["SCEN_PROB"].cpu().numpy() / total * pp]))
                    queue.append(k)
        for n in self.get_end_set():
            while n.parent_key is not None:
                dur = n.data.data["END_T"] - n.data.data["CUR_T"]
                dn = dt.get_node(n.key)
                if len(dn.data) == 1:
                    dn.data += [n.data.data["TRAJS_POS_HIST"][:, self.obs_len:self.obs_len + dur].cpu().numpy(),
                                n.data.data["TRAJS_COV_HIST"][:, self.obs_len:self.obs_len + dur].cpu().numpy(),
                                n.data.data["TGT_PTS"].cpu().numpy()]
                n = self.tree.get_node(n.parent_key)
        trees = []
        for key in dt.get_root().children_keys:
            st = Tree()
            n = dt.get_node(key)
            st.add_node(Node(n.key, None, n.data))
            queue = [n]
            while queue:
                c = queue.pop(0)
                for ck in c.children_keys:
                    cn = dt.get_node(ck)
                    st.add_node(Node(cn.key, c.key, cn.data))
                    queue.append(cn)
            trees.append(st)
        return trees


def flatten_trees(trees):
    """{node key: (parent, prob, trajs, covs, tgt_pts)} over a list of scenario trees."""
    out = {}
    for t in trees:
        for k, n in t.nodes.items():
            out[k] = (n.parent_key, float(n.data[0]), n.data[1], n.data[2], n.data[3])
    return out
