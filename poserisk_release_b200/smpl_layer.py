"""Drop-in for ``smplpytorch.pytorch.smpl_layer.SMPL_Layer``
(lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:14-158 of the reference).

Same constructor, buffers and ``forward(th_pose_axisang, th_betas, th_trans)``
contract; the arithmetic runs in libposerisk_b200.so (tcgen05 blend GEMM + fused
Rodrigues/chain and skinning kernels).  Outputs come back on the inputs' device,
float32, metres, exactly as the reference returns them.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.nn import Module

from . import _lib, _runtime
from .model_provider import GENDER_FILE, SMPLModelData, get_model_data


# (buffer name, field of SMPLModelData, leading batch axis): the seven tensors the reference layer registers
# (smpl_layer.py:40-56) -- their names are API, main/run.py and lib/utils/smpl.py read th_faces / th_J_regressor.
_BUFFERS = (('th_betas', 'betas', True), ('th_shapedirs', 'shapedirs', False), ('th_posedirs', 'posedirs', False),
            ('th_v_template', 'v_template', True), ('th_J_regressor', 'J_regressor', False),
            ('th_weights', 'weights', False))


class SMPL_Layer(Module):
    __constants__ = ['kintree_parents', 'gender', 'center_idx', 'num_joints']

    def __init__(self, center_idx=None, gender='neutral', model_root='smpl/native/models',
                 model_data: SMPLModelData | None = None, allow_synthetic: bool | None = None,
                 aligned_verts: bool = False):
        """center_idx: joint whose position is subtracted from vertices and joints when no translation is
        given (None: nothing is subtracted); gender: 'neutral' | 'female' | 'male', selects
        ``model_root/SMPL_<GENDER>.pkl``.  Extensions: ``model_data`` hands the constants over directly,
        ``allow_synthetic`` (or PRK_SYNTHETIC_SMPL=1) permits the synthetic stand-in when the licensed file
        is missing -- otherwise that is a FileNotFoundError, as in the reference; ``aligned_verts=True`` returns the
        vertices as a (B, 6890, 3) view over rows padded to 16 bytes (same values and indexing, not `is_contiguous()`),
        the layout the vertex kernel writes fastest (default: the reference's dense tensor)."""
        super().__init__()
        self.center_idx, self.gender = center_idx, gender
        self.aligned_verts = bool(aligned_verts)
        if gender in GENDER_FILE:
            self.model_path = os.path.join(model_root, GENDER_FILE[gender])
        if model_data is None:
            model_data = get_model_data(gender, model_root, allow_synthetic)
        self.smpl_data = model_data
        for name, field, batched in _BUFFERS:
            t = torch.tensor(np.asarray(getattr(model_data, field), dtype=np.float32))
            self.register_buffer(name, t.unsqueeze(0) if batched else t)
        self.register_buffer('th_faces', torch.from_numpy(np.asarray(model_data.faces).astype(np.int64)))
        self.vertice_segmentation = self.th_weights.argmax(dim=1)     # dominant joint per vertex
        self.kintree_table = model_data.kintree_table
        self.kintree_parents = [int(p) for p in self.kintree_table[0]]
        self.num_joints = len(self.kintree_parents)
        self._handles = {}

    # -- device-side model ---------------------------------------------------
    def _handle(self, device: torch.device) -> _runtime.ModelHandle:
        h = self._handles.get(device.index)
        if h is None:
            h = _runtime.ModelHandle(self.smpl_data, device)
            self._handles[device.index] = h
        return h

    @staticmethod
    def _optional(t, B, width, device, name):
        """None / the reference's ``torch.zeros(1)`` default -> None; else a (B, width) tensor."""
        if t is None:
            return None
        if t.numel() != B * width:
            # the reference evaluates torch.norm(t) == 0 first (smpl_layer.py:87,148); a
            # non-zero tensor of the wrong shape then fails inside matmul / broadcasting
            if t.numel() == 0 or bool(torch.norm(t.float()) == 0):
                return None
            raise RuntimeError(f'{name} must have shape ({B}, {width}), got {tuple(t.shape)}')
        return t.to(device=device, dtype=torch.float32).reshape(B, width).contiguous()

    def forward(self, th_pose_axisang, th_betas=torch.zeros(1), th_trans=torch.zeros(1), want_verts=True):
        """th_pose_axisang (B, 72): axis-angle pose, root first.  th_betas (B, 10): per-frame shape; the
        default / an all-zero batch selects the model's own betas.  th_trans (B, 3): added to vertices and
        joints; the default / an all-zero batch means no translation (then ``center_idx`` applies).
        want_verts=False (extension) skips the mesh and returns (None, joints).
        Returns (vertices (B, 6890, 3), joints (B, 24, 3)) in metres on the input's device."""
        out_device = th_pose_axisang.device
        device = _runtime.require_cuda(out_device)
        batch_size = th_pose_axisang.shape[0]
        pose = th_pose_axisang.to(device=device, dtype=torch.float32).reshape(batch_size, 72).contiguous()
        betas = self._optional(th_betas, batch_size, 10, device, 'th_betas')
        trans = self._optional(th_trans, batch_size, 3, device, 'th_trans')

        h = self._handle(device)
        L = _lib.lib()
        with torch.cuda.device(device):
            verts = None
            if want_verts:
                if self.aligned_verts and h.max_weights <= 4:
                    verts = _runtime.aligned_verts(batch_size, device)
                else:
                    verts = torch.empty((batch_size, 6890, 3), dtype=torch.float32, device=device)
            joints = torch.empty((batch_size, 24, 3), dtype=torch.float32, device=device)
            if batch_size > 0:
                ws, ws_bytes, _keep = _runtime.workspace.get(device, h.workspace_bytes(batch_size, not want_verts))
                _lib.check(L.prk_smpl_forward(
                    h.handle, _runtime.ptr(pose), _runtime.ptr(betas), _runtime.ptr(trans),
                    -1 if self.center_idx is None else int(self.center_idx), batch_size,
                    _runtime.ptr(verts), _runtime.verts_pitch(verts), _runtime.ptr(joints), ws, ws_bytes, _runtime.stream_ptr(device)))
        if out_device != device:
            joints = joints.to(out_device)
            verts = verts.to(out_device) if verts is not None else None
        return verts, joints
