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
    """One growable, 1024-byte aligned scratch buffer per (device, stream): calls queued on different streams
    may run at the same time and must not share the A' / A_j scratch or the host-path staging area."""

    def __init__(self):
        self._buf = {}
        self._lock = threading.Lock()

    def get(self, device: torch.device, nbytes: int):
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
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


class PeerComm:
    """Owns one prk_comm: this rank's gather buffer for the peer-memory all-gather (csrc/prk_comm.cu).

    `exchange_handles(blob) -> list of blobs in rank order` is the only thing the host transport has to do
    (torch.distributed.all_gather_object, MPI, ...)."""

    def __init__(self, rank: int, world: int, device: torch.device, slot_bytes: int, exchange_handles=None):
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.prk_comm_create(C.byref(h), rank, world, device.index, slot_bytes))
        self.handle = h
        self.rank, self.world, self.device, self.slot_bytes = rank, world, device, int(slot_bytes)
        self._finalizer = weakref.finalize(self, L.prk_comm_destroy, h)
        if world > 1:
            if exchange_handles is None:
                raise ValueError('exchange_handles is required when world > 1')
            self.open(exchange_handles(self.handle_blob()))

    def handle_blob(self) -> bytes:
        L = _lib.lib()
        buf = C.create_string_buffer(int(L.prk_comm_handle_bytes()))
        _lib.check(L.prk_comm_get_handle(self.handle, buf))
        return buf.raw

    def open(self, blobs):
        joined = b''.join(blobs)
        _lib.check(_lib.lib().prk_comm_open_peers(self.handle, joined))

    def allgather(self, local: torch.Tensor, row_offset: int, n_total_rows: int) -> torch.Tensor:
        """local: (n_local, ...) contiguous CUDA tensor whose rows belong at `row_offset`; returns a view of the
        gathered (n_total_rows, ...) array in this rank's buffer (valid until the next-but-one exchange)."""
        local = local.contiguous()
        row_shape = tuple(local.shape[1:])
        row_bytes = local.element_size() * int(np.prod(row_shape, dtype=np.int64)) if row_shape else local.element_size()
        out = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().prk_allgather_rows(self.handle, ptr(local) if local.numel() else None, local.shape[0],
                                                     row_offset, row_bytes, C.byref(out), stream_ptr(self.device)))
        return self.view(n_total_rows, row_shape, local.dtype, out.value)

    def gathered(self, n_rows: int, row_shape, dtype) -> torch.Tensor:
        """View of the slot written by the most recent exchange (e.g. one issued inside prk_pipeline, which runs it
        on a side stream): the current stream is made to wait for that exchange first."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().prk_comm_wait(self.handle, stream_ptr(self.device)))
        return self.view(n_rows, tuple(row_shape), dtype, _lib.lib().prk_comm_gathered(self.handle))

    def view(self, n_rows, row_shape, dtype, address):
        n_el = int(n_rows) * (int(np.prod(row_shape, dtype=np.int64)) if row_shape else 1)
        return _DevView(address, n_el, dtype, self.device, self).tensor.view((int(n_rows),) + tuple(row_shape))

    def check(self):
        _lib.check(_lib.lib().prk_comm_status(self.handle))


class _DevView:
    """torch tensor over library-owned device memory (the __cuda_array_interface__ route; `owner` keeps it alive)."""

    _TYPESTR = {torch.uint8: '|u1', torch.float32: '<f4', torch.float64: '<f8', torch.int32: '<i4', torch.int64: '<i8'}

    def __init__(self, address, n_el, dtype, device, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {'shape': (max(int(n_el), 0),), 'typestr': self._TYPESTR[dtype],
                                         'data': (int(address or 0), False), 'version': 2}
        if n_el > 0:
            with torch.cuda.device(device):
                self.tensor = torch.as_tensor(self, device=device)
        else:
            self.tensor = torch.empty(0, dtype=dtype, device=device)


def verts_pitch(verts) -> int:
    """Row pitch (floats) of a (B, 6890, 3) vertex tensor: 20670 when dense; a tensor whose frames are further apart
    (e.g. made by `aligned_verts`) must keep every frame's 20670 floats contiguous."""
    if verts is None:
        return 0
    if verts.dim() != 3 or tuple(verts.shape[1:]) != (6890, 3) or verts.dtype != torch.float32:
        raise ValueError('verts_out must be a float32 tensor of shape (B, 6890, 3)')
    if verts.shape[0] <= 1 and verts.is_contiguous():
        return 20670
    if verts.stride(2) != 1 or verts.stride(1) != 3:
        raise ValueError('verts_out: the 6890 x 3 floats of a frame must be contiguous')
    return 20670 if verts.shape[0] <= 1 else int(verts.stride(0))


def aligned_verts(B: int, device: torch.device) -> torch.Tensor:
    """(B, 6890, 3) float32 view over storage with 16-byte aligned rows (pitch 20672 floats): the layout the vertex
    kernel stores with bulk tensor (TMA) stores.  Values and indexing are those of a dense tensor; only
    `.is_contiguous()` differs (`.contiguous()` gives the reference's dense layout)."""
    flat = torch.empty((max(B, 1), _lib.VERTS_PITCH_ALIGNED), dtype=torch.float32, device=device)
    return flat[:B, :20670].view(B, 6890, 3)
