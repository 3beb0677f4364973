"""Host-side tensor plumbing of the predictor path (mirror of the torch half of the reference's
planners/mind/utils.py): ragged batching and frame helpers.  Pure tensor bookkeeping, no model
arithmetic; written for this repo (the reference functions are cited per function).
"""
import copy
from typing import Any, Dict, List

import torch


def to_device(data, device):
    """utils.py:9-20 (gpu): recursive, non-blocking transfer of every tensor."""
    if isinstance(data, (list, tuple)):
        return [to_device(x, device) for x in data]
    if isinstance(data, dict):
        return {k: to_device(v, device) for k, v in data.items()}
    if isinstance(data, torch.Tensor):
        return data.contiguous().to(device, non_blocking=True)
    return data


def pairwise_rpe(ctrs: torch.Tensor, vecs: torch.Tensor, radius: float = 100.0) -> torch.Tensor:
    """utils.py:193-242 (get_rpe): [cos a1, sin a1, cos a2, sin a2, 2*dist/radius] -> [5,M,M];
    entry [., a, b] relates vecs[b] to vecs[a] and to the displacement ctrs[b]-ctrs[a]."""
    d = ctrs.unsqueeze(0) - ctrs.unsqueeze(1)
    dist = d.norm(dim=-1)
    vb = vecs.unsqueeze(0).expand_as(d)
    va = vecs.unsqueeze(1).expand_as(d)
    nb, na = vb.norm(dim=-1), va.norm(dim=-1)
    den1 = nb * na + 1e-10
    den2 = nb * dist + 1e-10
    c1 = (vb[..., 0] * va[..., 0] + vb[..., 1] * va[..., 1]) / den1
    s1 = (vb[..., 0] * va[..., 1] - vb[..., 1] * va[..., 0]) / den1
    c2 = (vb[..., 0] * d[..., 0] + vb[..., 1] * d[..., 1]) / den2
    s2 = (vb[..., 0] * d[..., 1] - vb[..., 1] * d[..., 0]) / den2
    return torch.stack([c1, s1, c2, s2, dist * 2 / radius])


def origin_rotation(traj_pos, traj_ang, obs_len: int = 50):
    """utils.py:180-190: frame of the last observed step: orig [2], rot [[c,-s],[s,c]], theta."""
    orig = traj_pos[obs_len - 1]
    theta = traj_ang[obs_len - 1]
    c, s = torch.cos(theta), torch.sin(theta)
    rot = torch.stack([torch.stack([c, -s]), torch.stack([s, c])]).to(traj_pos.device)
    return orig, rot, theta


def actor_features(trajs: List[dict]):
    """utils.py:114-139 (actor_gather): [disp(2) | cos,sin heading(2) | vel(2) | type one-hot(7) |
    pad flag(1)] -> [sumNa, 14, 48] (first two steps dropped, :132) + arange index lists."""
    feats, idcs, count = [], [], 0
    for t in trajs:
        pos = t["TRAJS_POS_OBS"]
        disp = torch.zeros_like(pos)
        disp[:, 1:] = pos[:, 1:] - pos[:, :-1]
        f = torch.cat([disp, t["TRAJS_ANG_OBS"], t["TRAJS_VEL_OBS"], t["TRAJS_TYPE"], t["PAD_OBS"].unsqueeze(-1)], dim=-1)
        feats.append(f.transpose(1, 2))
        n = pos.shape[0]
        idcs.append(torch.arange(count, count + n))
        count += n
    return torch.cat(feats, 0)[..., 2:], idcs


def lane_features(graphs: List[dict]):
    """utils.py:75-111 (graph_gather): per-node [ctr(2) vec(2) intersect lane_type(3) cross_l(3)
    cross_r(3) left right] -> [sumNl, 10, 16] + index lists."""
    idcs, count, rows = [], 0, []
    for g in graphs:
        n = int(g["num_lanes"])
        idcs.append(torch.arange(count, count + n))
        count += n
        rows.append(torch.cat([g["node_ctrs"], g["node_vecs"], g["intersect"].unsqueeze(2), g["lane_type"],
                               g["cross_left"], g["cross_right"], g["left"].unsqueeze(2), g["right"].unsqueeze(2)], dim=-1))
    return torch.cat(rows, 0), idcs


def collate_scenes(batch: List[Dict[str, Any]]) -> Dict[str, Any]:
    """utils.py:142-168 (collate_fn): every key becomes a per-scene list; the seven network tensors
    are added (ACTORS, ACTOR_IDCS, LANES, LANE_IDCS, TGT_NODES stacked, TGT_RPE flattened to [B,20])."""
    data = {"BATCH_SIZE": len(batch)}
    for key in batch[0].keys():
        data[key] = [x[key] for x in batch]
    data["ACTORS"], data["ACTOR_IDCS"] = actor_features(data["TRAJS"])
    data["LANES"], data["LANE_IDCS"] = lane_features(data["LANE_GRAPH"])
    data["TGT_NODES"] = torch.stack(data["TGT_NODES"], 0)
    data["TGT_RPE"] = torch.stack(data["TGT_RPE"], 0).reshape(len(batch), -1)
    return data


def high_level_command(target_lane, target_lane_info, orig, rot, cur_vel, tar_time_ahead: float, min_vel: float = 0.5):
    """scenario_tree.py:613-652: 11 target-lane points about `tar_time_ahead` seconds ahead of the
    closest point, expressed as one 10-node polyline in its own instance frame.
    Returns tgt_pts [11,2] (global), tgt_nodes [10,16], (anchor pos [2], anchor dir [2])."""
    n = len(target_lane)
    closest = int(torch.argmin(torch.norm(target_lane - orig, dim=-1)))
    travel = max(float(cur_vel), min_vel) * tar_time_ahead
    idx = closest
    while idx < n - 1 and travel > 0:
        idx += 1
        travel -= float(torch.norm(target_lane[idx] - target_lane[idx - 1]))
    if idx == n - 1:
        idx -= 1
    idx = max(5, min(idx, n - 6))
    sel = torch.arange(idx - 5, idx + 6, device=target_lane.device)
    pts = target_lane[sel]
    info = target_lane_info[sel][1:]
    loc = torch.matmul(pts - orig, rot)
    anch_pos = loc.mean(dim=0)
    anch_vec = (loc[-1] - loc[0]) / torch.norm(loc[-1] - loc[0])
    anch_rot = torch.stack([torch.stack([anch_vec[0], -anch_vec[1]]), torch.stack([anch_vec[1], anch_vec[0]])])
    loc = torch.matmul(loc - anch_pos, anch_rot)
    ctrs = (loc[:-1] + loc[1:]) / 2.0
    vecs = loc[1:] - loc[:-1]
    return pts.clone(), torch.cat([ctrs, vecs, info], dim=-1), [anch_pos, anch_vec]


def pack_target_lane_info(info):
    """scenario_tree.py:110-120 (set_target_lane): 6 numpy arrays -> [M,12] tensor."""
    import numpy as np
    f = lambda a: torch.from_numpy(np.asarray(a))
    return torch.cat([f(info[0]).unsqueeze(1), f(info[1]), f(info[2]), f(info[3]), f(info[4]).unsqueeze(1),
                      f(info[5]).unsqueeze(1)], dim=-1)
