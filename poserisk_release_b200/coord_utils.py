"""Drop-ins for the hot-path functions of the reference's ``lib/utils/coord_utils.py``.

  get_joint_cam              coord_utils.py:7-21   (SMPL joints in mm, pelvis-relative)
  rot_to_angle               coord_utils.py:24-30  (cv2.Rodrigues: rotation matrices -> axis-angle)
  axis_angle_to_euler_angle  coord_utils.py:83-95  (cv2.Rodrigues + Euler XYZ, degrees)
All are batched on the GPU instead of looping per frame in Python; the results are
what the per-frame loops of the reference produce.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _runtime


def axis_angle_to_euler_angle(pose):
    """pose: (..., 3) axis-angle array (float32 or float64), typically (24, 3) for one frame
    or (N, 24, 3) for a batch.  Returns float64 degrees with the same shape, columns [x, y, z].

    Like the reference, a float32 pose goes through a float32 rotation matrix
    (cv2.Rodrigues keeps the input dtype) and a non-finite rotation fails the
    ``assert(isRotationMatrix(R))`` (coord_utils.py:70)."""
    if isinstance(pose, torch.Tensor):
        t = pose
        if t.dtype not in (torch.float32, torch.float64):
            raise TypeError('pose must be float32 or float64')
        device = _runtime.require_cuda(t.device)
        t = t.to(device).contiguous()
        shape = tuple(t.shape)
        return_numpy = False
    else:
        arr = np.asarray(pose)
        if arr.dtype not in (np.float32, np.float64):
            raise TypeError('pose must be float32 or float64 (cv2.Rodrigues accepts nothing else)')
        device = _runtime.require_cuda(None)
        shape = arr.shape
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
        return_numpy = True
    if len(shape) == 0 or shape[-1] != 3:
        raise ValueError('pose must have a trailing dimension of 3')
    n = t.numel() // 3
    out = torch.empty(shape, dtype=torch.float64, device=device)
    bad = torch.zeros(max(n, 1), dtype=torch.uint8, device=device)
    if n > 0:
        with torch.cuda.device(device):
            _lib.check(_lib.lib().prk_euler(
                _runtime.ptr(t), _lib.PRK_DTYPE_F32 if t.dtype == torch.float32 else _lib.PRK_DTYPE_F64,
                n, _runtime.ptr(out), _runtime.ptr(bad), _runtime.stream_ptr(device)))
        assert not bool(bad.any()), 'isRotationMatrix(R) failed'   # coord_utils.py:70
    return out.cpu().numpy() if return_numpy else out


def rot_to_angle(rotmat):
    """rotmat: (..., 3, 3) rotation matrices, float32 or float64 -- one frame's (24, 3, 3) as in the
    reference (base.py:226) or a whole batch (N, 24, 3, 3).  Returns the rotation vectors (..., 3) in
    the input's dtype: a numpy array for numpy input (what the reference returns), a CUDA tensor for
    tensor input (SPIN's pred_rotmat can stay on the device).  Like cv2.Rodrigues, a matrix that is
    not exactly orthonormal is first replaced by the nearest orthogonal matrix."""
    if isinstance(rotmat, torch.Tensor):
        t = rotmat
        if t.dtype not in (torch.float32, torch.float64):
            raise TypeError('rotmat must be float32 or float64')
        device = _runtime.require_cuda(t.device)
        t = t.to(device).contiguous()
        return_numpy = False
    else:
        arr = np.asarray(rotmat)
        if arr.dtype not in (np.float32, np.float64):
            raise TypeError('rotmat must be float32 or float64 (cv2.Rodrigues accepts nothing else)')
        device = _runtime.require_cuda(None)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
        return_numpy = True
    shape = tuple(t.shape)
    if len(shape) < 2 or shape[-2:] != (3, 3):
        raise ValueError('rotmat must have trailing dimensions (3, 3)')
    n = t.numel() // 9
    out = torch.empty(shape[:-2] + (3,), dtype=t.dtype, device=device)
    if n > 0:
        with torch.cuda.device(device):
            _lib.check(_lib.lib().prk_rot_to_angle(
                _runtime.ptr(t), _lib.PRK_DTYPE_F32 if t.dtype == torch.float32 else _lib.PRK_DTYPE_F64,
                n, _runtime.ptr(out), None, _runtime.stream_ptr(device)))
    return out.cpu().numpy() if return_numpy else out


def get_joint_cam(poses, smpl_model):
    """poses: (N, 24, 3) axis-angle array; smpl_model: object with ``layer['neutral']``.

    As in the reference every pose's root rotation is overwritten IN PLACE with
    [3.14, 0, 0] (coord_utils.py:10,13), betas are zero, the neutral layer is used, and
    the joints come back in millimetres relative to the pelvis, float32 (N, 24, 3)."""
    init_pose = np.array([3.14, 0, 0], dtype=np.float32)
    n = len(poses)
    if n == 0:
        raise ValueError('need at least one array to stack')     # np.stack([]) in the reference
    if isinstance(poses, np.ndarray):
        poses[:, 0] = init_pose
        batch = torch.from_numpy(np.ascontiguousarray(poses, dtype=np.float32).reshape(n, -1))
    else:
        for pose in poses:
            pose[0] = torch.tensor(init_pose) if isinstance(pose, torch.Tensor) else init_pose
        batch = torch.stack([torch.as_tensor(np.asarray(p), dtype=torch.float32).reshape(-1) for p in poses])
    layer = smpl_model.layer['neutral']
    device = _runtime.require_cuda(None)
    _, joints = layer(batch.to(device), torch.zeros((1, 10)), want_verts=False)
    joints = joints * 1000
    joints = joints - joints[:, :1]
    return joints.cpu().numpy()
