"""Seeded synthetic scene generators (S1, S2, ragged variant).

Pure tensor plumbing; produces exactly the 7-tuple ScenePredNet.pre_process
returns (reference planners/mind/networks/network.py:597-606):
  (ACTORS [sumNa,14,48], ACTOR_IDCS list, LANES [sumNl,10,16], LANE_IDCS list,
   RPE list of {'scene': [5,M,M], 'scene_mask': None}, TGT_NODES [B,10,16], TGT_RPE [B,20])
"""
import math
from typing import List, Optional, Sequence, Tuple

import torch

TWO_PI = 6.283185307179586


def pairwise_rpe(ctrs: torch.Tensor, vecs: torch.Tensor, radius: float = 100.0) -> torch.Tensor:
    """Host-side mirror of get_rpe (reference planners/mind/utils.py:193-212):
    5-channel relative encoding [cos a1, sin a1, cos a2, sin a2, 2*dist/radius], [5,M,M]."""
    d = ctrs.unsqueeze(0) - ctrs.unsqueeze(1)
    dist = d.norm(dim=-1)
    vb = vecs.unsqueeze(0).expand_as(d)
    va = vecs.unsqueeze(1).expand_as(d)
    nb, na, nd = vb.norm(dim=-1), va.norm(dim=-1), dist
    den1 = nb * na + 1e-10
    den2 = nb * nd + 1e-10
    c1 = (vb[..., 0] * va[..., 0] + vb[..., 1] * va[..., 1]) / den1
    s1 = (vb[..., 0] * va[..., 1] - vb[..., 1] * va[..., 0]) / den1
    c2 = (vb[..., 0] * d[..., 0] + vb[..., 1] * d[..., 1]) / den2
    s2 = (vb[..., 0] * d[..., 1] - vb[..., 1] * d[..., 0]) / den2
    return torch.stack([c1, s1, c2, s2, dist * 2 / radius])


def scene_s1(seed: int = 1234, n_actor: int = 32, n_lane: int = 128, with_geom: bool = False):
    """One scene, draw order fixed by SURVEY.md 8d 'S1'."""
    g = torch.Generator().manual_seed(seed)
    actors = torch.randn(n_actor, 14, 48, generator=g)
    lanes = torch.randn(n_lane, 10, 16, generator=g)
    m = n_actor + n_lane
    ctrs = torch.randn(m, 2, generator=g) * 30
    th = torch.rand(m, generator=g) * TWO_PI
    vecs = torch.stack([torch.cos(th), torch.sin(th)], dim=-1)
    rpe = pairwise_rpe(ctrs, vecs)
    tgt_nodes = torch.randn(1, 10, 16, generator=g)
    tgt_rpe = torch.randn(1, 20, generator=g)
    out = dict(actors=actors, lanes=lanes, rpe=rpe, tgt_nodes=tgt_nodes, tgt_rpe=tgt_rpe)
    if with_geom:
        out["ctrs"], out["vecs"] = ctrs, vecs
    return out


def batch_from_scenes(scenes: Sequence[dict]):
    """Ragged concatenation + index lists, as collate_fn does
    (reference planners/mind/utils.py:114-168)."""
    actors = torch.cat([s["actors"] for s in scenes], 0)
    lanes = torch.cat([s["lanes"] for s in scenes], 0)
    a_idcs, l_idcs, ca, cl = [], [], 0, 0
    for s in scenes:
        na, nl = s["actors"].shape[0], s["lanes"].shape[0]
        a_idcs.append(torch.arange(ca, ca + na))
        l_idcs.append(torch.arange(cl, cl + nl))
        ca += na
        cl += nl
    rpe = [{"scene": s["rpe"], "scene_mask": None} for s in scenes]
    tgt_nodes = torch.cat([s["tgt_nodes"] for s in scenes], 0)
    tgt_rpe = torch.cat([s["tgt_rpe"] for s in scenes], 0)
    return actors, a_idcs, lanes, l_idcs, rpe, tgt_nodes, tgt_rpe


def batch_s2(batch: int = 256, n_actor: int = 32, n_lane: int = 128, seed0: int = 1000):
    """Config-2 throughput batch: `batch` S1-style scenes with seeds seed0+b."""
    return batch_from_scenes([scene_s1(seed0 + b, n_actor, n_lane) for b in range(batch)])


def batch_ragged(batch: int = 8, seed: int = 7, na_rng=(8, 32), nl_rng=(32, 128), seed0: int = 2000):
    """Ragged variant: Na ~ U{8..32}, Nl ~ U{32..128} (SURVEY.md 8d S2 note)."""
    g = torch.Generator().manual_seed(seed)
    scenes = []
    for b in range(batch):
        na = int(torch.randint(na_rng[0], na_rng[1] + 1, (1,), generator=g))
        nl = int(torch.randint(nl_rng[0], nl_rng[1] + 1, (1,), generator=g))
        scenes.append(scene_s1(seed0 + b, na, nl))
    return batch_from_scenes(scenes)


def random_state_dict(seed: int = 0, like: Optional[dict] = None):
    """Seeded random weights with the reference's 328 keys/shapes (used when the
    shipped checkpoint is not wanted).  `like` supplies key->shape."""
    assert like is not None
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in like.items():
        shape = tuple(v.shape) if hasattr(v, "shape") else tuple(v)
        if len(shape) == 1:
            if k.endswith("weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                sd[k] = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[k] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
    return sd


# ---------------------------------------------------------------------------------------------
# S3: kinematic synthetic scene for the AIME tree.  Restates what
# process_data (reference scenario_tree.py:136-206) builds, on hand-made inputs (no av2/shapely).
# ---------------------------------------------------------------------------------------------
def scene_s3(x0=(50, 60, 40, 80, 75, 65), y0=(0, 3.5, -3.5, 3.5, 0, -3.5), v=(6, 5, 7, 4, 6, 5),
             tar_time_ahead: float = 5.0):
    """Returns (collated one-scene dict, target_lane [301,2] float32, target_lane_info list of 6 arrays,
    lane_graph dict).  Actor 0 is the ego ('AV')."""
    import numpy as np
    from . import plumbing as P
    na = len(x0)
    t = torch.arange(50, dtype=torch.float32) * 0.1
    pos = torch.stack([torch.stack([x0[i] + v[i] * (t - 4.9), torch.full_like(t, float(y0[i]))], -1) for i in range(na)])
    ang = torch.zeros(na, 50)
    vel = torch.stack([torch.stack([torch.full_like(t, float(v[i])), torch.zeros_like(t)], -1) for i in range(na)])
    ttype = torch.zeros(na, 50, 7, dtype=torch.int16)
    ttype[..., 0] = 1
    pad = torch.ones(na, 50, dtype=torch.int16)
    # target lane: y = 0, x = 0..300 @ 1 m
    lane = np.stack([np.arange(301, dtype=np.float32), np.zeros(301, dtype=np.float32)], -1)
    one = lambda k: np.tile(np.eye(3, dtype=np.float32)[k], (301, 1))
    info = [np.zeros(301, np.float32), one(0), one(0), one(1), np.ones(301, np.float32), np.zeros(301, np.float32)]
    tl = torch.from_numpy(lane)
    tli = P.pack_target_lane_info(info)

    orig, rot, theta = P.origin_rotation(pos[0], ang[0])
    # lane graph: 3 lanes x 20 segments of 15 m, 11 points each, scene frame then instance frame
    node_ctrs, node_vecs, lane_ctrs, lane_vecs = [], [], [], []
    for y in (-3.5, 0.0, 3.5):
        for s in range(20):
            xs = torch.linspace(15.0 * s, 15.0 * (s + 1), 11)
            p = torch.stack([xs, torch.full_like(xs, y)], -1)
            p = torch.matmul(p - orig, rot)
            anch = p.mean(0)
            d = (p[-1] - p[0]) / torch.norm(p[-1] - p[0])
            r = torch.stack([torch.stack([d[0], -d[1]]), torch.stack([d[1], d[0]])])
            q = torch.matmul(p - anch, r)
            node_ctrs.append((q[:-1] + q[1:]) / 2.0)
            node_vecs.append(q[1:] - q[:-1])
            lane_ctrs.append(anch)
            lane_vecs.append(d)
    nl = 60
    oh = lambda k: torch.tensor([1 if i == k else 0 for i in range(3)], dtype=torch.int16).repeat(nl, 10, 1)
    graph = dict(node_ctrs=torch.stack(node_ctrs), node_vecs=torch.stack(node_vecs),
                 intersect=torch.zeros(nl, 10, dtype=torch.int16), lane_type=oh(0), cross_left=oh(0), cross_right=oh(0),
                 left=torch.ones(nl, 10, dtype=torch.int16), right=torch.ones(nl, 10, dtype=torch.int16),
                 lane_ctrs=torch.stack(lane_ctrs), lane_vecs=torch.stack(lane_vecs), num_nodes=nl * 10, num_lanes=nl)

    # scene-normalise, then per-actor normalise (scenario_tree.py:136-158)
    spos = torch.matmul(pos - orig, rot)
    sang = ang - theta
    svel = torch.matmul(vel, rot)
    pn, an, vn, ctrs, vecs = [], [], [], [], []
    for i in range(na):
        o, r, th = P.origin_rotation(spos[i], sang[i])
        pn.append(torch.matmul(spos[i] - o, r))
        an.append(sang[i] - th)
        vn.append(torch.matmul(svel[i], r))
        ctrs.append(o)
        vecs.append(torch.stack([torch.cos(th), torch.sin(th)]))
    an = torch.stack(an)
    trajs = dict(TRAJS_POS_OBS=torch.stack(pn), TRAJS_ANG_OBS=torch.stack([torch.cos(an), torch.sin(an)], -1),
                 TRAJS_VEL_OBS=torch.stack(vn), TRAJS_TYPE=ttype, PAD_OBS=pad, TRAJS_CTRS=torch.stack(ctrs),
                 TRAJS_VECS=torch.stack(vecs), TRAJS_TID=["AV"] + [str(i) for i in range(1, na)],
                 TRAJS_CAT=["av"] + ["exo"] * (na - 1))
    tgt_pts, tgt_nodes, tgt_anch = P.high_level_command(tl, tli, orig, rot, float(v[0]), tar_time_ahead)
    rpe = {"scene": P.pairwise_rpe(torch.cat([trajs["TRAJS_CTRS"], graph["lane_ctrs"]]),
                                   torch.cat([trajs["TRAJS_VECS"], graph["lane_vecs"]])), "scene_mask": None}
    tgt_rpe = P.pairwise_rpe(torch.stack([tgt_anch[0], trajs["TRAJS_CTRS"][0]]), torch.stack([tgt_anch[1], trajs["TRAJS_VECS"][0]]))
    data = dict(ORIG=orig, ROT=rot, TRAJS=trajs, LANE_GRAPH=graph, TGT_PTS=tgt_pts, TGT_NODES=tgt_nodes, TGT_ANCH=tgt_anch,
                RPE=rpe, TGT_RPE=tgt_rpe)
    return P.collate_scenes([data]), lane, info, graph
