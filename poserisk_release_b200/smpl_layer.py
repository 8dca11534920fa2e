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


class SMPL_Layer(Module):
    __constants__ = ['kintree_parents', 'gender', 'center_idx', 'num_joints']

    def __init__(self, center_idx=None, gender='neutral', model_root='smpl/native/models',
                 model_data: SMPLModelData | None = None):
        """
        Args:
            center_idx: index of center joint in our computations,
            model_root: path to pkl files for the model
            gender: 'neutral' (default) or 'female' or 'male'
            model_data: (extension) pre-loaded constants; default loads
                ``model_root/SMPL_<GENDER>.pkl`` or, when that licensed file is absent,
                the synthetic SMPL-shaped model.
        """
        super().__init__()
        self.center_idx = center_idx
        self.gender = gender
        if gender in GENDER_FILE:                      # smpl_layer.py:30-35
            self.model_path = os.path.join(model_root, GENDER_FILE[gender])
        data = model_data if model_data is not None else get_model_data(gender, model_root)
        self.smpl_data = data

        self.register_buffer('th_betas', torch.Tensor(data.betas).unsqueeze(0))
        self.register_buffer('th_shapedirs', torch.Tensor(data.shapedirs))
        self.register_buffer('th_posedirs', torch.Tensor(data.posedirs))
        self.register_buffer('th_v_template', torch.Tensor(data.v_template).unsqueeze(0))
        self.register_buffer('th_J_regressor', torch.Tensor(np.array(data.J_regressor)))
        self.register_buffer('th_weights', torch.Tensor(data.weights))
        self.register_buffer('th_faces', torch.Tensor(data.faces.astype(np.int32)).long())

        self.vertice_segmentation = torch.argmax(self.th_weights, dim=1)

        # Kinematic chain params
        self.kintree_table = data.kintree_table
        parents = list(self.kintree_table[0].tolist())
        self.kintree_parents = parents
        self.num_joints = len(parents)  # 24
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
        """
        Args:
        th_pose_axisang (Tensor (batch_size x 72)): pose parameters in axis-angle representation
        th_betas (Tensor (batch_size x 10)): if provided, uses given shape parameters
        th_trans (Tensor (batch_size x 3)): if provided, applies trans to joints and vertices
        want_verts (extension): False selects the joints-only fast path and returns (None, joints)
        """
        out_device = th_pose_axisang.device
        device = _runtime.require_cuda(out_device)
        batch_size = th_pose_axisang.shape[0]
        pose = th_pose_axisang.to(device=device, dtype=torch.float32).reshape(batch_size, 72).contiguous()
        betas = self._optional(th_betas, batch_size, 10, device, 'th_betas')
        trans = self._optional(th_trans, batch_size, 3, device, 'th_trans')

        h = self._handle(device)
        L = _lib.lib()
        with torch.cuda.device(device):
            verts = torch.empty((batch_size, 6890, 3), dtype=torch.float32, device=device) if want_verts else None
            joints = torch.empty((batch_size, 24, 3), dtype=torch.float32, device=device)
            if batch_size > 0:
                ws, ws_bytes, _keep = _runtime.workspace.get(device, h.workspace_bytes(batch_size, not want_verts))
                _lib.check(L.prk_smpl_forward(
                    h.handle, _runtime.ptr(pose), _runtime.ptr(betas), _runtime.ptr(trans),
                    -1 if self.center_idx is None else int(self.center_idx), batch_size,
                    _runtime.ptr(verts), _runtime.ptr(joints), ws, ws_bytes, _runtime.stream_ptr(device)))
        if out_device != device:
            joints = joints.to(out_device)
            verts = verts.to(out_device) if verts is not None else None
        # Vertices and joints in meters
        return verts, joints
