"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

Frames are independent: every rank runs the whole path on a contiguous
range of frames with its own replica of the model constants, and the only exchange is
one all-gather of the 32-byte per-frame score records and, optionally, of the debug
Euler sequences (the reference's per-frame result lists, lib/core/base.py:144-151,168).
Vertices and joints stay rank-local.

Two transports:
  * peer memory (default on GPUs): ``prk_allgather_rows`` of libposerisk_b200.so -- every rank stores its
    rows straight into every rank's gather buffer over NVLink (csrc/prk_comm.cu); torch.distributed only
    hands the IPC handles round once.  The engine can issue it underneath the body-model kernels.
  * NCCL ``all_gather_into_tensor`` (``transport='nccl'``, and the automatic fall-back when CUDA IPC is not
    permitted in the container); gloo runs the same code in the CPU tests of this module's logic.
"""
from __future__ import annotations

import numpy as np
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


def shard_tracks(track_lengths, world: int):
    """Whole tracks per rank (BASELINE.json config 4: 16 tracks / 8 GPUs = 2 each): contiguous blocks of
    tracks, the first n % world ranks get one extra.  Returns per rank (first_track, last_track + 1,
    first_frame, last_frame + 1) with frames counted in track order."""
    n = len(track_lengths)
    starts = np.concatenate(([0], np.cumsum(track_lengths))).astype(np.int64)
    out = []
    for r in range(world):
        t0, t1 = shard_range(n, r, world)
        out.append((t0, t1, int(starts[t0]), int(starts[t1])))
    return out


def all_gather_rows(local: torch.Tensor, n_total: int, group=None, sizes=None) -> torch.Tensor:
    """Gathers per-frame rows (first dim = this rank's frames) into frame order on every rank with the
    process group's collective.  `sizes` = rows per rank (default: shard_sizes).  Equal shards use one
    all_gather_into_tensor; ragged shards are padded to the largest shard first."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if sizes is None:
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


class ScoreExchange:
    """All-gather of score records (N,32) uint8 and debug Euler rows (N,k,3) float64 for `n_total` frames.

    transport: 'peer' (prk_allgather_rows over peer memory), 'nccl' (the process group's collective), or 'auto'
    (peer memory when every rank could map its peers, else the collective).  `used` says which one runs."""

    def __init__(self, n_total: int, device=None, n_debug: int = 0, group=None, transport: str = 'auto'):
        self.n_total, self.n_debug, self.group = int(n_total), int(n_debug), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device
        self.comm_scores = self.comm_euler = None
        self.used = 'nccl' if self.world > 1 else 'none'
        self.why_not_peer = None
        if transport in ('auto', 'peer') and device is not None and torch.device(device).type == 'cuda':
            self._open_peer(transport == 'peer')

    def _open_peer(self, required: bool):
        from . import _lib, _runtime
        dev = torch.device(self.device)

        def exchange(blob: bytes):
            if self.world == 1:
                return [blob]
            blobs = [None] * self.world
            dist.all_gather_object(blobs, blob, group=self.group)
            return blobs

        err = None
        try:
            cs = _runtime.PeerComm(self.rank, self.world, dev, max(self.n_total, 1) * 32, exchange)
            ce = None
            if self.n_debug:
                ce = _runtime.PeerComm(self.rank, self.world, dev, max(self.n_total, 1) * self.n_debug * 24, exchange)
        except _lib.PeerExchangeError as e:      # e.g. cudaIpcOpenMemHandle refused in this container
            err = str(e)
            cs = ce = None
        # every rank must take the same decision
        if self.world > 1:
            flags = [None] * self.world
            dist.all_gather_object(flags, err, group=self.group)
            err = next((f for f in flags if f), None)
        if err is None:
            self.comm_scores, self.comm_euler, self.used = cs, ce, 'peer'
        else:
            self.why_not_peer = err
            if required:
                raise _lib.PeerExchangeError(err)

    # -- used by PoseRiskEngine.run(exchange=...) ---------------------------------------------
    def native_handles(self, n_debug: int):
        """(prk_comm* for the records, prk_comm* for the Euler rows) or (None, None) on the NCCL transport."""
        if self.used != 'peer':
            return None, None
        return self.comm_scores.handle, (self.comm_euler.handle if (n_debug and self.comm_euler is not None) else None)

    def collect(self, local_scores, frame_offset: int, local_euler=None, sizes=None):
        """Gathered (scores, euler|None) on every rank.  On the peer transport after an engine call that was given
        this exchange the rows are already there (views of the gather buffer); otherwise the exchange happens here."""
        if self.used == 'peer':
            s = self.comm_scores.gathered(self.n_total, (32,), torch.uint8)
            e = None
            if local_euler is not None and self.comm_euler is not None:
                e = self.comm_euler.gathered(self.n_total, (self.n_debug, 3), torch.float64)
            return s, e
        s = all_gather_rows(local_scores, self.n_total, self.group, sizes)
        e = None if local_euler is None else all_gather_rows(local_euler, self.n_total, self.group, sizes)
        return s, e

    def gather(self, local_scores, frame_offset: int, local_euler=None, sizes=None):
        """Stand-alone exchange (the rows were produced by calls that did not take part in it, e.g. config 4's
        per-gender runs): peer stores or the collective, on the current stream."""
        if self.used == 'peer':
            s = self.comm_scores.allgather(local_scores, frame_offset, self.n_total)
            e = None
            if local_euler is not None and self.comm_euler is not None:
                e = self.comm_euler.allgather(local_euler, frame_offset, self.n_total)
            return s, e
        return self.collect(local_scores, frame_offset, local_euler, sizes)

    def join(self):
        """Make the current stream wait for the most recent exchange issued inside an engine call (peer transport: it
        runs on side streams and is not joined by the call; NCCL runs on the current stream already)."""
        if self.used == 'peer':
            from . import _lib, _runtime
            for c in (self.comm_scores, self.comm_euler):
                if c is not None:
                    with torch.cuda.device(c.device):
                        _lib.check(_lib.lib().prk_comm_wait(c.handle, _runtime.stream_ptr(c.device)))

    def check(self):
        if self.comm_scores is not None:
            self.comm_scores.check()
        if self.comm_euler is not None:
            self.comm_euler.check()


def run_sharded(engine, pose, betas, trans, add_info, track_of_frame=None, want_verts=False, group=None,
                debug_joints=None, exchange=None, transport='auto'):
    """Every rank passes the FULL (host or device) inputs; each processes its contiguous shard and the score
    records -- and the debug Euler sequences when `debug_joints` is given -- are all-gathered.  Returns
    dict(scores (N,32) uint8 and euler (N,k,3) float64 on every rank, joints / verts of the local shard, range)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = pose.shape[0]
    lo, hi = shard_range(n, rank, world)
    sl = slice(lo, hi)
    n_debug = 0 if debug_joints is None else len(debug_joints)
    if exchange is None:
        exchange = ScoreExchange(n, getattr(engine, 'device', None), n_debug, group, transport)
    out = engine.run(pose[sl], None if betas is None else betas[sl], None if trans is None else trans[sl],
                     add_info=add_info, track_of_frame=None if track_of_frame is None else track_of_frame[sl],
                     want_verts=want_verts, debug_joints=debug_joints, exchange=exchange, frame_offset=lo)
    scores, euler = exchange.collect(out['scores'], lo, out.get('euler'))
    return {'scores': scores, 'euler': euler, 'joints': out['joints'], 'verts': out['verts'], 'range': (lo, hi),
            'transport': exchange.used}
