"""Data-parallel sharding of independent scenes / tree branches over ranks (one process per GPU,
torch.distributed: NCCL on GPUs, gloo in CPU tests).  The forward has no communication; the only
collective is ONE all-gather of the decoded outputs (cls, reg, vel) at the end of a batch
(the reference itself is single-device, so this is new functionality)."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of scenes for `rank`: sizes differ by at most one, order preserved."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_scenes(scenes: Sequence, rank: int, world: int) -> List:
    s, e = shard_range(len(scenes), rank, world)
    return list(scenes[s:e])


def all_gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather along dim 0 of tensors whose dim-0 length may differ per rank (ragged actor counts):
    lengths are exchanged first, shards padded to the maximum, one all_gather_into_tensor, then cropped.
    Every rank ends with the concatenation in rank order."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = torch.empty(world, device=t.device, dtype=torch.int64)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = [int(x) for x in sizes.tolist()]
    mx = max(sizes)
    if all(s == mx for s in sizes):
        out = torch.empty((world * mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[: t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], 0)


def all_gather_predictions(cls: torch.Tensor, reg: torch.Tensor, vel: torch.Tensor, group=None):
    """The path's single collective: every rank ends with all scenes' (cls [B,6], reg [A,6,60,5],
    vel [A,6,60,2]) in scene order."""
    return all_gather_rows(cls, group), all_gather_rows(reg, group), all_gather_rows(vel, group)


def all_gather_packed(pack: torch.Tensor, B: int, A: int, group=None, out: torch.Tensor = None, async_op: bool = False):
    """ONE collective for a batch with equal shards (every rank B scenes / A actors: the benchmark batch, a tree level):
    `pack` is forward_packed's output buffer cls | reg | vel (predictor.pack_layout), written by the decoder kernels
    themselves, so nothing is packed or copied before the send.  Returns (gathered [world, len(pack)], work handle);
    `unpack_gathered` gives per-tensor views of it."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, pack.numel()), device=pack.device, dtype=pack.dtype)
    work = dist.all_gather_into_tensor(out.view(-1), pack, group=group, async_op=async_op)
    return out, work


def unpack_gathered(gathered: torch.Tensor, B: int, A: int):
    """(cls [world, B, 6], reg [world, A, 6, 60, 5], vel [world, A, 6, 60, 2]) as strided views of the gathered buffer:
    rank r's scenes are [r]; flattening the first two dims (scene order) costs one copy per tensor, only if needed."""
    from .predictor import pack_layout
    (c0, cn), (r0, rn), (v0, vn) = pack_layout(B, A)
    w = gathered.shape[0]
    return (gathered[:, c0:c0 + cn].view(w, B, 6), gathered[:, r0:r0 + rn].view(w, A, 6, 60, 5),
            gathered[:, v0:v0 + vn].view(w, A, 6, 60, 2))


# ---- tree mode: the frontier of one depth level sharded over ranks ------------------------------
def shard_level_inputs(net_in, geom, n_frontier: int, rank: int, world: int):
    """Contiguous block of frontier scenes for `rank`.  `net_in` is the level's network input tuple
    (actors [F*Na,14,48], actor index lists, lanes [F*Nl,10,16], lane index lists, rpe | None, tgt_nodes [F,10,16],
    tgt_rpe [F,20]); all scenes of a level have the same actor / lane counts (they are copies of one root scene), so
    the shard is a set of views: no copy.  `geom` = (ctrs, vecs) [F*(Na+Nl), 2] or None.  Returns (net_in, geom, (s, e))."""
    actors, a_idcs, lanes, l_idcs, rpe, tgt_nodes, tgt_rpe = net_in[:7]
    F = n_frontier
    s, e = shard_range(F, rank, world)
    na, nl = actors.shape[0] // F, lanes.shape[0] // F
    if actors.shape[0] != F * na or lanes.shape[0] != F * nl:
        raise ValueError("level inputs are not uniform over the frontier")
    sub = (actors[s * na:e * na], list(a_idcs[s:e]), lanes[s * nl:e * nl], list(l_idcs[s:e]),
           None if rpe is None else list(rpe[s:e]), tgt_nodes[s:e], tgt_rpe[s:e])
    g = None
    if geom is not None:
        m = geom[0].shape[0] // F
        g = (geom[0][s * m:e * m], geom[1][s * m:e * m])
    return sub, g, (s, e)


def sharded_level_forward(forward, net_in, geom, n_frontier: int, group=None):
    """One depth level of the scenario tree on all ranks: every rank predicts its block of the frontier with
    `forward(net_in, geom) -> (cls [f,6], reg [f*Na,6,60,5], vel [f*Na,6,60,2], ...)` and ONE all-gather of the decoded
    outputs leaves every rank with the whole level, in frontier order.  The tree bookkeeping that follows (prune / merge /
    branch decisions) is replicated: all ranks see identical gathered tensors and therefore build identical trees.
    Falls back to a replicated forward when the frontier is smaller than the world (root, natural trees)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1 or n_frontier < world:
        out = forward(net_in, geom)
        return out[0], out[1], out[2]
    sub, g, (s0, e0) = shard_level_inputs(net_in, geom, n_frontier, dist.get_rank(group), world)
    out = forward(sub, g)
    pack = out[6] if len(out) > 6 else None
    if pack is not None and n_frontier % world == 0:
        # equal shards: one all-gather of the packed cls | reg | vel buffer the decoder wrote
        f, a = e0 - s0, out[1].shape[0]
        gathered, _ = all_gather_packed(pack, f, a, group)
        c, r, v = unpack_gathered(gathered, f, a)
        return c.reshape(-1, 6), r.reshape((-1,) + tuple(r.shape[2:])), v.reshape((-1,) + tuple(v.shape[2:]))
    return all_gather_predictions(out[0], out[1], out[2], group)
