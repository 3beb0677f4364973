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
