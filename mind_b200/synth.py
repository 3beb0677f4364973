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
