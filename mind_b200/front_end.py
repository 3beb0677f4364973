"""The scene front end of the scenario-tree generator: observation tracks + static map -> the collated one-scene dict
the predictor and the tree rollout consume (the step immediately before the hot path).

Restates, on the host (numpy fp64 / torch fp32, the same arithmetic order as the reference so that the result is
bit-identical), `ScenarioTreeGenerator.process_data` (planners/mind/scenario_tree.py:122-206) with its helpers
`get_agent_trajectories` (planners/mind/utils.py:245-342) and `update_lane_graph_from_argo` (utils.py:345-483).
The map-dependent half of the lane graph (arc-length resampling of every centreline into ~15 m polylines of 11 points
and the per-polyline attribute rows) does not depend on the ego pose: it is computed once per map and cached; the
reference redoes it in every plan call (10 Hz).

Needs `av2` / `shapely` API objects; when the real packages are absent `mind_b200.compat` provides them.
"""
import copy
from typing import Dict, List

import numpy as np
import torch

from . import plumbing as P

OBS_LEN = 50
NUM_SEG_POINTS = 10
SEG_LENGTH = 15.0

_TYPE_SLOT = {"vehicle": 0, "pedestrian": 1, "motorcyclist": 2, "cyclist": 3, "bus": 4, "unknown": 5}   # else 6 (utils.py:298-311)
_CROSSABLE = {"DASH_SOLID_YELLOW", "DASH_SOLID_WHITE", "DASHED_WHITE", "DASHED_YELLOW", "DOUBLE_DASH_YELLOW", "DOUBLE_DASH_WHITE"}
_NOT_CROSSABLE = {"DOUBLE_SOLID_YELLOW", "DOUBLE_SOLID_WHITE", "SOLID_YELLOW", "SOLID_WHITE", "SOLID_DASH_WHITE",
                  "SOLID_DASH_YELLOW", "SOLID_BLUE"}
_LANE_SLOT = {"VEHICLE": 0, "BIKE": 1, "BUS": 2}


def _val(e):
    return getattr(e, "value", e)


def _fill_nearest(values: np.ndarray, seen: np.ndarray) -> np.ndarray:
    """common/data.py:24-44 (padding_traj_nn): unobserved steps take the previous observed value, leading ones the first."""
    idx = np.where(seen, np.arange(len(seen)), -1)
    prev = np.maximum.accumulate(idx)
    first = int(np.argmax(seen))
    return values[np.where(prev >= 0, prev, first)]


def agent_trajectories(agent_obs) -> Dict[str, object]:
    """utils.py:245-342.  agent_obs: dict track_id -> Track (object_states of the last <= 50 steps, `observed` flags).
    Order: 'AV' first, then the other tracks in dict order; tracks not observed at the last step are dropped."""
    keys = list(agent_obs.keys())
    if "AV" not in keys:
        raise KeyError("agent_obs has no 'AV' track")
    order = ["AV"] + [k for k in keys if k != "AV"]
    pos, ang, vel, typ, flags, tids, cats = [], [], [], [], [], [], []
    for key in order:
        st = agent_obs[key].object_states
        if st[-1].observed is False:
            continue
        n = len(st)
        # one pass over the states: (observed, x, y, heading, vx, vy); unobserved steps contribute zeros / no type row
        rows = np.array([(s.observed, s.position[0], s.position[1], s.heading, s.velocity[0], s.velocity[1]) for s in st],
                        dtype=np.float64)
        seen50 = np.zeros(OBS_LEN, dtype=bool)
        seen50[OBS_LEN - n:] = rows[:, 0] != 0
        full = np.zeros((OBS_LEN, 5))
        full[OBS_LEN - n:] = np.where(rows[:, :1] != 0, rows[:, 1:], 0.0)
        one_hot = np.zeros(7)
        one_hot[_TYPE_SLOT.get(str(_val(agent_obs[key].object_type)).lower(), 6)] = 1
        t50 = np.zeros((OBS_LEN, 7))
        t50[seen50] = one_hot
        pos.append(_fill_nearest(full[:, 0:2], seen50))
        ang.append(_fill_nearest(full[:, 2], seen50))
        vel.append(full[:, 3:5])
        typ.append(t50)
        flags.append(seen50.astype(np.int64))
        tids.append(key)
        cats.append("av" if key == "AV" else "exo")
    f32 = lambda x: torch.from_numpy(np.array(x).astype(np.float32))
    i16 = lambda x: torch.from_numpy(np.array(x).astype(np.int16))
    return dict(pos=f32(pos), ang=f32(ang), vel=f32(vel), type=i16(typ), flags=i16(flags), tid=tids, cat=cats)


class _MapPolylines:
    """pose-independent half of update_lane_graph_from_argo (utils.py:345-371,393-452), cached per static map"""

    def __init__(self, static_map):
        from shapely.geometry import LineString
        self.pts: List[np.ndarray] = []         # per polyline: [11, 2] fp64 points in the map frame
        rows = dict(lane_type=[], intersect=[], cross_left=[], cross_right=[], left=[], right=[])
        for lane_id, lane in static_map.vector_lane_segments.items():
            cl_raw = static_map.get_lane_segment_centerline(lane_id)[:, 0:2]
            assert cl_raw.shape[0] == NUM_SEG_POINTS, "[Error] Wrong num of points in lane - {}:{}".format(lane_id, cl_raw.shape[0])
            ls = LineString(cl_raw)
            num_segs = np.max([int(np.floor(ls.length / SEG_LENGTH)), 1])
            ds = ls.length / num_segs
            lt = np.zeros(3)
            lt[_LANE_SLOT[str(_val(lane.lane_type))]] = 1            # KeyError = "[Error] Wrong lane type"

            def cross(mark):
                m, c = str(_val(mark)), np.zeros(3)
                c[0 if m in _CROSSABLE else 1 if m in _NOT_CROSSABLE else 2] = 1
                return c
            for i in range(num_segs):
                cl_pts = [ls.interpolate(s) for s in np.linspace(i * ds, (i + 1) * ds, NUM_SEG_POINTS + 1)]
                self.pts.append(np.array(LineString(cl_pts).coords))
                rows["lane_type"].append(np.tile(lt, (NUM_SEG_POINTS, 1)))
                rows["intersect"].append(np.full(NUM_SEG_POINTS, 1.0 if lane.is_intersection else 0.0, np.float32))
                rows["cross_left"].append(np.tile(cross(lane.left_mark_type), (NUM_SEG_POINTS, 1)))
                rows["cross_right"].append(np.tile(cross(lane.right_mark_type), (NUM_SEG_POINTS, 1)))
                rows["left"].append(np.full(NUM_SEG_POINTS, 0.0 if lane.left_neighbor_id is None else 1.0, np.float32))
                rows["right"].append(np.full(NUM_SEG_POINTS, 0.0 if lane.right_neighbor_id is None else 1.0, np.float32))
        self.attrs = {k: np.stack(v, axis=0).astype(np.int16) for k, v in rows.items()}

    def lane_graph(self, orig: np.ndarray, rot: np.ndarray) -> dict:
        """pose-dependent half (utils.py:372-391,454-483): scene frame, per-polyline instance frame"""
        node_ctrs, node_vecs, lane_ctrs, lane_vecs = [], [], [], []
        for pts in self.pts:
            ctrln = (pts - orig).dot(rot)
            anch_pos = np.mean(ctrln, axis=0)
            anch_vec = (ctrln[-1] - ctrln[0]) / np.linalg.norm(ctrln[-1] - ctrln[0])
            anch_rot = np.array([[anch_vec[0], -anch_vec[1]], [anch_vec[1], anch_vec[0]]])
            lane_ctrs.append(anch_pos)
            lane_vecs.append(anch_vec)
            ctrln = (ctrln - anch_pos).dot(anch_rot)
            node_ctrs.append(np.asarray((ctrln[:-1] + ctrln[1:]) / 2.0, np.float32))
            node_vecs.append(np.asarray(ctrln[1:] - ctrln[:-1], np.float32))
        g = dict(node_ctrs=np.stack(node_ctrs, axis=0).astype(np.float32), node_vecs=np.stack(node_vecs, axis=0).astype(np.float32),
                 lane_ctrs=np.array(lane_ctrs).astype(np.float32), lane_vecs=np.array(lane_vecs).astype(np.float32))
        g = {k: torch.from_numpy(v) for k, v in g.items()}
        for k in ("lane_type", "intersect", "cross_left", "cross_right", "left", "right"):
            g[k] = torch.from_numpy(self.attrs[k].copy())
        g["num_nodes"] = g["node_ctrs"].shape[0] * g["node_ctrs"].shape[1]
        g["num_lanes"] = g["lane_ctrs"].shape[0]
        return g


class ArgoFrontEnd:
    """Callable (lcl_smp, agent_obs) -> collated scene dict on the generator's device; `generator` supplies the target
    lane (set_target_lane), the tree configuration and receives `lane_graph` (scenario_tree.py:204)."""

    def __init__(self, generator, device=None):
        self.gen = generator
        self.device = torch.device(device) if device is not None else generator.device
        self._maps = {}

    def polylines(self, static_map) -> _MapPolylines:
        key = id(static_map)
        hit = self._maps.get(key)
        if hit is None or hit[0] is not static_map:
            if len(self._maps) > 8:
                self._maps.clear()
            hit = self._maps[key] = (static_map, _MapPolylines(static_map))
        return hit[1]

    def __call__(self, lcl_smp, agent_obs):
        from . import compat
        compat.install()
        tr = agent_trajectories(agent_obs)
        pos, ang, vel = tr["pos"], tr["ang"], tr["vel"]
        cur_vel = lcl_smp.ego_agent.state[2]
        orig, rot, theta = P.origin_rotation(pos[0], ang[0])                                 # :128 (target-centric)
        graph = self.polylines(lcl_smp.map_data).lane_graph(orig.numpy(), rot.numpy())       # :131
        pos = torch.matmul(pos - orig, rot)                                                  # :136-138
        ang = ang - theta
        vel = torch.matmul(vel, rot)
        pn, an, vn, ctrs, vecs = [], [], [], [], []
        for p, a, v in zip(pos, ang, vel):                                                   # :146-152
            o, r, th = P.origin_rotation(p, a)
            pn.append(torch.matmul(p - o, r))
            an.append(a - th)
            vn.append(torch.matmul(v, r))
            ctrs.append(o)
            vecs.append(torch.stack([torch.cos(th), torch.sin(th)]))
        an = torch.stack(an)
        trajs = dict(TRAJS_POS_OBS=torch.stack(pn), TRAJS_ANG_OBS=torch.stack([torch.cos(an), torch.sin(an)], dim=-1),
                     TRAJS_VEL_OBS=torch.stack(vn), TRAJS_TYPE=tr["type"], PAD_OBS=tr["flags"],
                     TRAJS_CTRS=torch.stack(ctrs), TRAJS_VECS=torch.stack(vecs), TRAJS_TID=tr["tid"], TRAJS_CAT=tr["cat"])
        gen = self.gen
        tgt_pts, tgt_nodes, anch = P.high_level_command(gen.target_lane.detach().cpu(), gen.target_lane_info.detach().cpu(),
                                                        orig, rot, cur_vel, gen.config.tar_time_ahead)     # :176
        rpe = {"scene": P.pairwise_rpe(torch.cat([trajs["TRAJS_CTRS"], graph["lane_ctrs"]], 0),
                                       torch.cat([trajs["TRAJS_VECS"], graph["lane_vecs"]], 0)), "scene_mask": None}
        tgt_rpe = P.pairwise_rpe(torch.stack([anch[0], trajs["TRAJS_CTRS"][0]]), torch.stack([anch[1], trajs["TRAJS_VECS"][0]]))
        data = dict(ORIG=orig, ROT=rot, TRAJS=trajs, LANE_GRAPH=graph, TGT_PTS=tgt_pts, TGT_NODES=tgt_nodes, TGT_ANCH=anch,
                    RPE=rpe, TGT_RPE=tgt_rpe)
        gen.lane_graph = copy.deepcopy(graph)                                                # :204
        return P.to_device(P.collate_scenes([data]), self.device)
