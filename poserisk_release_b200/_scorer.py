"""Shared plumbing of the REBA / RULA drop-ins: Euler degrees -> device scorer -> the
reference's per-frame result dicts."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _runtime

JOINT_NAME = ('Pelvis', 'L_Hip', 'R_Hip', 'Torso', 'L_Knee', 'R_Knee', 'Spine', 'L_Ankle', 'R_Ankle', 'Chest',
              'L_Toe', 'R_Toe', 'Neck', 'L_Thorax', 'R_Thorax', 'Head', 'L_Shoulder', 'R_Shoulder', 'L_Elbow', 'R_Elbow',
              'L_Wrist', 'R_Wrist', 'L_Hand', 'R_Hand')


def score_euler_records(poses, add_info, which: int, track_of_frame=None, device=None) -> np.ndarray:
    """poses: (N,24,3) Euler degrees (anything np.asarray accepts, or a CUDA float64 tensor).
    Returns a structured numpy array of prk_score_rec."""
    device = _runtime.require_cuda(device)
    if isinstance(poses, torch.Tensor):
        e = poses.to(device=device, dtype=torch.float64).reshape(-1, 24, 3).contiguous()
    else:
        arr = np.asarray(poses, dtype=np.float64)
        if arr.size == 0:
            arr = arr.reshape(0, 24, 3)
        e = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1, 24, 3))).to(device)
    n = e.shape[0]
    info = _runtime.addinfo_tensor(add_info, device)
    track = None
    if track_of_frame is not None:
        track = torch.as_tensor(track_of_frame, dtype=torch.int32).to(device).contiguous()
    out = torch.zeros((n, 32), dtype=torch.uint8, device=device)
    if n > 0:
        with torch.cuda.device(device):
            _lib.check(_lib.lib().prk_score_euler(_runtime.ptr(e), _runtime.ptr(info), info.shape[0], _runtime.ptr(track), n,
                                                  which, _runtime.ptr(out), _runtime.stream_ptr(device)))
    return _runtime.records_to_numpy(out)


def action_band(bands, score):
    """(level, text) of the band a rounded score falls into; (None, None) below the first band, as the
    reference's if-ladders leave both unset for scores < 1 (reba.py:83-104, rula.py:100-118)."""
    score = round(score)
    if score < 1:
        return None, None
    for top, level, text in bands:
        if top is None or score <= top:
            return level, text
