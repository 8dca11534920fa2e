"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

Frames are independent: every rank runs the whole path on a contiguous
range of frames with its own replica of the model constants, and the only exchange is
one all-gather of the 32-byte per-frame score records (and, optionally, the debug
Euler sequences).  Vertices and joints stay rank-local.  Backend: NCCL over
NVLink/NVSwitch on GPUs; gloo is used by the CPU tests of this module's logic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`: the first n % world ranks get one extra frame."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_sizes(n_frames: int, world: int):
    return [shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gathers per-frame rows (first dim = this rank's frames, sharded with shard_range)
    into frame order on every rank.  Equal shards use one all_gather_into_tensor; ragged
    shards are padded to the largest shard first."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(n_total, world)
    mx = max(sizes)
    row_shape = tuple(local.shape[1:])
    if local.shape[0] != mx:
        pad = torch.zeros((mx,) + row_shape, dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        local = pad
    out = torch.empty((world * mx,) + row_shape, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    parts = [out[r * mx:r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def run_sharded(engine, pose, betas, trans, add_info, track_of_frame=None, want_verts=False, group=None):
    """Every rank passes the FULL (host or device) inputs; each processes its shard and the
    score records are all-gathered.  Returns dict(scores (N,32) uint8 on every rank,
    joints / verts of the local shard, range)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = pose.shape[0]
    lo, hi = shard_range(n, rank, world)
    sl = slice(lo, hi)
    out = engine.run(pose[sl], None if betas is None else betas[sl], None if trans is None else trans[sl],
                     add_info=add_info, track_of_frame=None if track_of_frame is None else track_of_frame[sl],
                     want_verts=want_verts)
    scores = all_gather_rows(out['scores'], n, group)
    return {'scores': scores, 'joints': out['joints'], 'verts': out['verts'], 'range': (lo, hi)}
