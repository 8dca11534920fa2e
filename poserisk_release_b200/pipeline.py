"""Batched frame engine: what ``lib/core/base.py:225-239,151,168`` does per frame in
Python (Euler angles, SMPL forward, REBA, RULA) done for whole batches on one GPU.

``PoseRiskEngine.run`` takes device tensors, ``run_host`` takes (pinned) host arrays
and goes through ``prk_pipeline_host`` (host->device copies, kernels, device->host
copies of scores and joints; back-to-back calls overlap their copies with the kernels).  Multi-person tracks with different
genders / additional information are handled by grouping frames per gender
(BASELINE.json config 4).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, _runtime
from .model_provider import get_model_data

GENDERS = ('neutral', 'female', 'male')


class PoseRiskEngine:
    def __init__(self, device=None, genders=('neutral',), model_root=None, allow_synthetic=None, model_data=None,
                 aligned_verts=True):
        """model_data: optional {gender: SMPLModelData}; otherwise ``model_root/SMPL_<GENDER>.pkl`` is read
        (FileNotFoundError when missing, unless allow_synthetic / PRK_SYNTHETIC_SMPL=1: model_provider.get_model_data).
        aligned_verts: vertex tensors the engine allocates itself get 16-byte aligned rows (`_runtime.aligned_verts`: same
        shape, values and indexing, row pitch 20672 instead of 20670 floats) so that the vertex kernel can store them with
        bulk tensor stores; a `verts_out` the caller passes is used as it is, dense or aligned."""
        self.aligned_verts = bool(aligned_verts)
        self.device = _runtime.require_cuda(device)
        self.models = {}
        for g in genders:
            data = (model_data or {}).get(g) or get_model_data(g, model_root, allow_synthetic)
            self.models[g] = _runtime.ModelHandle(data, self.device)
        self._host_ws = None

    # ------------------------------------------------------------------ device path
    def _dev_f32(self, t, B, n):
        """(B, n) float32 contiguous tensor on the engine's device (no copy when it already is one)."""
        if t is None:
            return None
        if not (t.device == self.device and t.dtype == torch.float32 and t.is_contiguous()):
            t = t.to(self.device, torch.float32).contiguous()
        return t.reshape(B, n)

    def run(self, pose, betas=None, trans=None, add_info=None, track_of_frame=None, gender='neutral',
            want_verts=True, center_idx=None, verts_out=None, joints_out=None, scores_out=None,
            debug_joints=None, euler_out=None, exchange=None, frame_offset=0):
        """pose (B,72) float32 CUDA tensor.  Returns dict(verts|None, joints, scores[, euler]) where
        scores is a (B,32) uint8 tensor of prk_score_rec.  The *_out tensors, when given, receive
        the results (no allocation in the call).
        debug_joints: joint ids whose Euler sequences (B,k,3) float64 are emitted as well (the
        --debug_joints output, base.py:144-146).
        exchange: a distributed.ScoreExchange -- the records (and Euler rows) of this rank are all-gathered
        into every rank's buffer at frame `frame_offset`, underneath the body-model kernels."""
        dev = self.device
        B = pose.shape[0]
        pose = self._dev_f32(pose, B, 72)
        betas = self._dev_f32(betas, B, 10)
        trans = self._dev_f32(trans, B, 3)
        info = add_info if isinstance(add_info, torch.Tensor) else _runtime.addinfo_tensor(add_info, dev)
        track = None if track_of_frame is None else torch.as_tensor(track_of_frame, dtype=torch.int32).to(dev).contiguous()
        ids = None if debug_joints is None else np.ascontiguousarray(debug_joints, np.int32)
        n_debug = 0 if ids is None else len(ids)
        h = self.models[gender]
        comm_s = comm_e = None
        if exchange is not None:
            comm_s, comm_e = exchange.native_handles(n_debug)
        with torch.cuda.device(dev):
            verts = None
            if want_verts:
                if verts_out is not None:
                    verts = verts_out
                elif self.aligned_verts and h.max_weights <= 4:
                    verts = _runtime.aligned_verts(B, dev)
                else:
                    verts = torch.empty((B, 6890, 3), dtype=torch.float32, device=dev)
            joints = joints_out if joints_out is not None else torch.empty((B, 24, 3), dtype=torch.float32, device=dev)
            scores = scores_out if scores_out is not None else torch.empty((B, 32), dtype=torch.uint8, device=dev)
            euler = None
            if n_debug:
                euler = euler_out if euler_out is not None else torch.empty((B, n_debug, 3), dtype=torch.float64, device=dev)
            if B > 0 or comm_s is not None:
                key = (B, want_verts, gender)
                if getattr(self, '_ws_key', None) != key:      # workspace size per (batch, mode): queried once
                    self._ws_key, self._ws_bytes = key, h.workspace_bytes(max(B, 1), not want_verts)
                ws, ws_bytes, _keep = _runtime.workspace.get(dev, self._ws_bytes)
                _lib.check(_lib.lib().prk_pipeline(
                    h.handle, _runtime.ptr(pose), _runtime.ptr(betas), _runtime.ptr(trans),
                    -1 if center_idx is None else int(center_idx), _runtime.ptr(info), info.shape[0], _runtime.ptr(track), B,
                    _runtime.ptr(verts), _runtime.verts_pitch(verts), _runtime.ptr(joints), _runtime.ptr(scores), _runtime.ptr(euler),
                    None if ids is None else ids.ctypes.data_as(C.c_void_p), n_debug, comm_s, comm_e, int(frame_offset),
                    ws, ws_bytes, _runtime.stream_ptr(dev)))
        out = {'verts': verts, 'joints': joints, 'scores': scores}
        if n_debug:
            out['euler'] = euler
        return out

    def run_tracks(self, pose, betas, trans, add_infos, track_of_frame, gender_of_track, want_verts=True,
                   verts_out=None, joints_out=None, scores_out=None):
        """Mixed-gender multi-person batch (BASELINE.json config 4).  Frames of one track are contiguous
        (a track is one person's clip), so the batch is a sequence of RUNS of equal gender: every run goes
        through its gender's model as a slice of the caller's tensors and writes straight into slices of the
        outputs -- no index gather, no scatter, no device synchronisation.  `track_of_frame` must be a HOST
        array (it is metadata the caller built; base.py:72-73 picks tracks on the host too)."""
        dev = self.device
        B = pose.shape[0]
        track_h = np.ascontiguousarray(track_of_frame.cpu().numpy() if isinstance(track_of_frame, torch.Tensor)
                                       else track_of_frame, dtype=np.int32).reshape(B)
        n_tracks = len(gender_of_track)
        if B and (track_h.min() < 0 or track_h.max() >= n_tracks):
            raise IndexError('track_of_frame holds an id outside gender_of_track')    # list index in the reference
        g_of_t = np.array([GENDERS.index(g) for g in gender_of_track], np.int32)
        g_of_f = g_of_t[track_h] if B else np.zeros(0, np.int32)
        cuts = np.flatnonzero(np.diff(g_of_f)) + 1                   # run boundaries
        starts = np.concatenate(([0], cuts)) if B else np.zeros(0, np.int64)
        ends = np.concatenate((cuts, [B])) if B else np.zeros(0, np.int64)
        pose = self._dev_f32(pose, B, 72)
        betas = self._dev_f32(betas, B, 10)
        trans = self._dev_f32(trans, B, 3)
        info = add_infos if isinstance(add_infos, torch.Tensor) else _runtime.addinfo_tensor(add_infos, dev)
        track_d = torch.from_numpy(track_h).to(dev)
        joints = joints_out if joints_out is not None else torch.empty((B, 24, 3), dtype=torch.float32, device=dev)
        scores = scores_out if scores_out is not None else torch.empty((B, 32), dtype=torch.uint8, device=dev)
        verts = None
        if want_verts:
            if verts_out is not None:
                verts = verts_out
            elif self.aligned_verts and all(self.models[g].max_weights <= 4 for g in set(gender_of_track)):
                verts = _runtime.aligned_verts(B, dev)
            else:
                verts = torch.empty((B, 6890, 3), dtype=torch.float32, device=dev)
        for lo, hi in zip(starts.tolist(), ends.tolist()):
            self.run(pose[lo:hi], None if betas is None else betas[lo:hi], None if trans is None else trans[lo:hi],
                     info, track_d[lo:hi], gender=GENDERS[int(g_of_f[lo])], want_verts=want_verts,
                     verts_out=None if verts is None else verts[lo:hi], joints_out=joints[lo:hi], scores_out=scores[lo:hi])
        return {'verts': verts, 'joints': joints, 'scores': scores, 'runs': len(starts)}

    def euler_debug(self, pose, joint_ids, add_info, track_of_frame=None):
        """Scores + Euler sequences of the listed joints (the --debug_joints output,
        base.py:144-146): returns (scores (B,32) uint8, euler (B,k,3) float64)."""
        dev = self.device
        B = pose.shape[0]
        is64 = pose.dtype == torch.float64
        pose = pose.to(dev).reshape(B, 72).contiguous()
        info = add_info if isinstance(add_info, torch.Tensor) else _runtime.addinfo_tensor(add_info, dev)
        track = None if track_of_frame is None else torch.as_tensor(track_of_frame, dtype=torch.int32).to(dev).contiguous()
        ids = np.asarray(joint_ids, np.int32)
        scores = torch.empty((B, 32), dtype=torch.uint8, device=dev)
        eul = torch.empty((B, len(ids), 3), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().prk_score_pose(
                _runtime.ptr(pose), _lib.PRK_DTYPE_F64 if is64 else _lib.PRK_DTYPE_F32, _runtime.ptr(info), info.shape[0],
                _runtime.ptr(track), B, _lib.PRK_SCORE_REBA | _lib.PRK_SCORE_RULA, _runtime.ptr(scores),
                _runtime.ptr(eul), ids.ctypes.data_as(C.c_void_p), len(ids), _runtime.stream_ptr(dev)))
        return scores, eul

    # ------------------------------------------------------------------ host path
    def run_host(self, pose, betas, trans, add_infos, track_of_frame, joints_out, scores_out, gender='neutral',
                 verts_out=None, center_idx=None, exchange=None, frame_offset=0):
        """Host buffers in, host buffers out (torch CPU tensors, ideally pinned):
        pose (B,72) f32, betas (B,10)|None, trans (B,3)|None, joints_out (B,24,3) f32,
        scores_out (B,32) uint8; verts_out is an optional CUDA tensor.  Asynchronous on the
        current stream: synchronize before reading the outputs."""
        dev = self.device
        B = pose.shape[0]
        h = self.models[gender]
        info = _lib.addinfo_array(add_infos)
        track = None if track_of_frame is None else np.ascontiguousarray(track_of_frame, np.int32)
        joints_only = verts_out is None
        with torch.cuda.device(dev):
            ws, ws_bytes, _keep = _runtime.host_workspace.get(dev, h.host_workspace_bytes(B, joints_only))
            _lib.check(_lib.lib().prk_pipeline_host(
                h.handle, _runtime.ptr(pose), _runtime.ptr(betas), _runtime.ptr(trans),
                -1 if center_idx is None else int(center_idx), info.ctypes.data_as(C.c_void_p), info.shape[0],
                None if track is None else track.ctypes.data_as(C.c_void_p), B, _runtime.ptr(verts_out),
                _runtime.verts_pitch(verts_out), _runtime.ptr(joints_out), _runtime.ptr(scores_out),
                None if exchange is None else exchange.native_handles(0)[0], int(frame_offset),
                ws, ws_bytes, _runtime.stream_ptr(dev)))
        self._keep = (info, track)   # pageable host arrays must outlive the async copies
        # device copy of the score records inside the workspace (what a multi-GPU caller all-gathers)
        base = ws.value - _keep.data_ptr()
        off = base + int(_lib.lib().prk_host_scores_offset(h.handle, B))
        self.host_scores_device = _keep[off:off + B * 32].view(B, 32)
        return joints_out, scores_out

    # ------------------------------------------------------------------ aggregation
    def aggregate(self, scores, which='REBA'):
        """Predictor.post_processing numbers (base.py:260-271) from a device histogram:
        (mean, top-50% mean, top-10% mean, max, mode), each rounded to 3 decimals like the
        reference (top-10% is nan when fewer than 10 frames)."""
        dev = self.device
        B = scores.shape[0]
        hist = torch.empty(64, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().prk_score_histogram(
                _runtime.ptr(scores), B, _lib.PRK_SCORE_REBA if which == 'REBA' else _lib.PRK_SCORE_RULA,
                _runtime.ptr(hist), _runtime.stream_ptr(dev)))
        return aggregate_from_histogram(hist.cpu().numpy(), B)


def aggregate_from_histogram(hist: np.ndarray, n: int):
    """hist[k] = number of frames with score k-16.  Mirrors base.py:260-271: scores sorted
    descending, mean, mean of first n//2, mean of first n//10, max, scipy mode (smallest of
    the most frequent values)."""
    values = np.arange(64) - 16

    def top_mean(k):
        if k == 0:
            return float('nan')
        left, tot = k, 0.0
        for v in range(63, -1, -1):
            take = min(left, int(hist[v]))
            tot += take * float(values[v])
            left -= take
            if left == 0:
                break
        return tot / k

    if n == 0:
        return (float('nan'),) * 5
    mean = float((hist * values).sum()) / n
    mx = int(values[np.nonzero(hist)[0].max()])
    mode = int(values[int(np.argmax(hist))])          # argmax returns the first = smallest value
    return (round(mean, 3), round(top_mean(n // 2), 3), round(top_mean(n // 10), 3), round(mx, 3), mode)
