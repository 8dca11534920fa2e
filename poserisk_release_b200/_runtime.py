"""Device plumbing shared by the drop-in classes: torch supplies device memory and
streams, the work itself is done by libposerisk_b200.so."""
from __future__ import annotations

import ctypes as C
import threading
import weakref

import numpy as np
import torch

from . import _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('poserisk_release_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    if device.type != 'cuda':
        return torch.device('cuda', torch.cuda.current_device())
    if device.index is None:
        return torch.device('cuda', torch.cuda.current_device())
    return device


def stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class _Workspace:
    """One growable, 1024-byte aligned scratch buffer per device."""

    def __init__(self):
        self._buf = {}
        self._lock = threading.Lock()

    def get(self, device: torch.device, nbytes: int):
        key = device.index
        with self._lock:
            b = self._buf.get(key)
            if b is None or b.numel() < nbytes + 1024:
                if b is not None:
                    # the library also works on its own streams, which the caching allocator does not
                    # know about: nothing may still be using the old buffer when it is released
                    torch.cuda.synchronize(device)
                b = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
                self._buf[key] = b
        base = b.data_ptr()
        off = (-base) % 1024
        return C.c_void_p(base + off), b.numel() - off, b


workspace = _Workspace()
host_workspace = _Workspace()     # staging + scratch of the host-buffer path, never shared with device-API calls


class ModelHandle:
    """Owns one prk_model (device copy of the packed model constants)."""

    def __init__(self, data, device: torch.device):
        L = _lib.lib()
        f32 = np.float32
        vt = np.ascontiguousarray(data.v_template, f32).reshape(6890, 3)
        sd = np.ascontiguousarray(data.shapedirs, f32).reshape(6890, 3, 10)
        pd = np.ascontiguousarray(data.posedirs, f32).reshape(6890, 3, 207)
        jr = np.ascontiguousarray(data.J_regressor, f32).reshape(24, 6890)
        w = np.ascontiguousarray(data.weights, f32).reshape(6890, 24)
        parents = np.array([-1] + [int(x) for x in list(data.parents)[1:]], np.int32)
        betas = np.ascontiguousarray(data.betas, f32).reshape(10)
        h = C.c_void_p()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(L.prk_model_create(C.byref(h), device.index, p(vt), p(sd), p(pd), p(jr), p(w), p(parents), p(betas)))
        self.handle = h
        self.device = device
        self.max_weights = int(L.prk_model_max_weights(h))
        self._finalizer = weakref.finalize(self, L.prk_model_destroy, h)

    def workspace_bytes(self, B: int, joints_only: bool) -> int:
        return int(_lib.lib().prk_workspace_bytes(self.handle, B, _lib.PRK_FLAG_JOINTS_ONLY if joints_only else 0))

    def host_workspace_bytes(self, B: int, joints_only: bool) -> int:
        return int(_lib.lib().prk_host_workspace_bytes(self.handle, B, _lib.PRK_FLAG_JOINTS_ONLY if joints_only else 0))


def records_to_numpy(rec_u8: torch.Tensor) -> np.ndarray:
    """(N, 32) uint8 tensor of prk_score_rec -> structured numpy array on the host."""
    return rec_u8.cpu().numpy().reshape(-1).view(_lib.REC_DTYPE)


def addinfo_tensor(add_infos, device: torch.device) -> torch.Tensor:
    return torch.from_numpy(_lib.addinfo_array(add_infos)).to(device)
