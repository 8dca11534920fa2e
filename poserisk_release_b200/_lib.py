"""ctypes binding of csrc/libposerisk_b200.so (C ABI: include/poserisk_b200.h).

There is no CPU or PyTorch fallback: if the library is missing the import fails
loudly (build it with ``python -m poserisk_release_b200.build`` or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB

PRK_OK = 0
PRK_SCORE_REBA = 1
PRK_SCORE_RULA = 2
PRK_DTYPE_F32 = 0
PRK_DTYPE_F64 = 1
PRK_FLAG_JOINTS_ONLY = 1
VERTS_PITCH_ALIGNED = 20672     # floats per vertex row with 16-byte aligned rows (PRK_VERTS_PITCH_ALIGNED)
PRK_ERR_PEER = 6
ABI_VERSION = 2

REBA_KEYS = ("Legs_bilateral_weight_bearing/walking", "Sitting", "Load/Force Score",
             "Arm_supported_leaning_L", "Arm_supported_leaning_R", "Coupling", "Activity_Score")
RULA_KEYS = ("Arm_supported_leaning_L", "Arm_supported_leaning_R", "A_Muscle_use_L",
             "A_Muscle_use_R", "A_Load/Force_L", "A_Load/Force_R",
             "Legs_bilateral_weight_bearing", "B_Muscle_use", "B_Load/Force")

# struct prk_score_rec (32 bytes)
REC_DTYPE = np.dtype([('reba_score', '<i2'), ('rula_score', '<i2'), ('reba_parts', 'u1', (9,)),
                      ('rula_parts', 'u1', (11,)), ('flags', 'u1'), ('pad', 'u1', (7,))])
assert REC_DTYPE.itemsize == 32

EXPORTS = (
    'prk_abi_version', 'prk_strerror', 'prk_last_error_detail', 'prk_model_create',
    'prk_model_destroy', 'prk_model_device', 'prk_model_max_weights', 'prk_workspace_bytes',
    'prk_smpl_forward', 'prk_score_pose', 'prk_score_euler', 'prk_euler', 'prk_pipeline',
    'prk_rot_to_angle', 'prk_host_workspace_bytes', 'prk_host_scores_offset', 'prk_pipeline_host', 'prk_score_histogram', 'prk_debug_blend', 'prk_debug_unit_range',
    'prk_vposed_pitch', 'prk_launch_count', 'prk_profile_begin', 'prk_profile_begin_stages', 'prk_profile_end',
    'prk_comm_create', 'prk_comm_destroy', 'prk_comm_handle_bytes', 'prk_comm_get_handle', 'prk_comm_open_peers',
    'prk_comm_gathered', 'prk_comm_wait', 'prk_allgather_rows', 'prk_allgather_scores', 'prk_comm_status')


class PoseRiskError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB):
        raise ImportError(
            f'{LIB} not found: the CUDA extension is required (no CPU fallback). '
            'Build it with `python -m poserisk_release_b200.build`.')
    L = C.CDLL(LIB)
    vp, i32, i64, u32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_size_t
    L.prk_abi_version.restype = i32
    L.prk_strerror.restype = C.c_char_p
    L.prk_strerror.argtypes = [i32]
    L.prk_last_error_detail.restype = C.c_char_p
    L.prk_model_create.restype = i32
    L.prk_model_create.argtypes = [C.POINTER(vp), i32, vp, vp, vp, vp, vp, vp, vp]
    L.prk_model_destroy.restype = None
    L.prk_model_destroy.argtypes = [vp]
    L.prk_model_device.restype = i32
    L.prk_model_device.argtypes = [vp]
    L.prk_model_max_weights.restype = i32
    L.prk_model_max_weights.argtypes = [vp]
    L.prk_workspace_bytes.restype = sz
    L.prk_workspace_bytes.argtypes = [vp, i64, u32]
    L.prk_smpl_forward.restype = i32
    L.prk_smpl_forward.argtypes = [vp, vp, vp, vp, i32, i64, vp, i64, vp, vp, sz, vp]
    L.prk_score_pose.restype = i32
    L.prk_score_pose.argtypes = [vp, i32, vp, i32, vp, i64, u32, vp, vp, vp, i32, vp]
    L.prk_score_euler.restype = i32
    L.prk_score_euler.argtypes = [vp, vp, i32, vp, i64, u32, vp, vp]
    L.prk_euler.restype = i32
    L.prk_euler.argtypes = [vp, i32, i64, vp, vp, vp]
    L.prk_pipeline.restype = i32
    L.prk_pipeline.argtypes = [vp, vp, vp, vp, i32, vp, i32, vp, i64, vp, i64, vp, vp, vp, vp, i32, vp, vp, i64, vp, sz, vp]
    L.prk_rot_to_angle.restype = i32
    L.prk_rot_to_angle.argtypes = [vp, i32, i64, vp, vp, vp]
    L.prk_host_workspace_bytes.restype = sz
    L.prk_host_workspace_bytes.argtypes = [vp, i64, u32]
    L.prk_host_scores_offset.restype = sz
    L.prk_host_scores_offset.argtypes = [vp, i64]
    L.prk_pipeline_host.restype = i32
    L.prk_pipeline_host.argtypes = [vp, vp, vp, vp, i32, vp, i32, vp, i64, vp, i64, vp, vp, vp, i64, vp, sz, vp]
    L.prk_score_histogram.restype = i32
    L.prk_score_histogram.argtypes = [vp, i64, u32, vp, vp]
    L.prk_debug_blend.restype = i32
    L.prk_debug_blend.argtypes = [vp, vp, vp, i64, vp, i32, vp, sz, vp]
    L.prk_debug_unit_range.restype = i32
    L.prk_debug_unit_range.argtypes = [i64, i64, i32, i64, vp, vp]
    L.prk_vposed_pitch.restype = i64
    L.prk_launch_count.restype = C.c_uint64
    L.prk_profile_begin.restype = i32
    L.prk_profile_begin_stages.restype = i32
    L.prk_profile_begin_stages.argtypes = [u32]
    L.prk_profile_end.restype = i32
    L.prk_profile_end.argtypes = [vp, vp]
    L.prk_comm_create.restype = i32
    L.prk_comm_create.argtypes = [C.POINTER(vp), i32, i32, i32, sz]
    L.prk_comm_destroy.restype = None
    L.prk_comm_destroy.argtypes = [vp]
    L.prk_comm_handle_bytes.restype = sz
    L.prk_comm_get_handle.restype = i32
    L.prk_comm_get_handle.argtypes = [vp, vp]
    L.prk_comm_open_peers.restype = i32
    L.prk_comm_open_peers.argtypes = [vp, vp]
    L.prk_comm_wait.restype = i32
    L.prk_comm_wait.argtypes = [vp, vp]
    L.prk_comm_gathered.restype = vp
    L.prk_comm_gathered.argtypes = [vp]
    L.prk_allgather_rows.restype = i32
    L.prk_allgather_rows.argtypes = [vp, vp, i64, i64, i64, C.POINTER(vp), vp]
    L.prk_allgather_scores.restype = i32
    L.prk_allgather_scores.argtypes = [vp, vp, i64, i64, C.POINTER(vp), vp]
    L.prk_comm_status.restype = i32
    L.prk_comm_status.argtypes = [vp]
    if L.prk_abi_version() != ABI_VERSION:
        raise ImportError(f'libposerisk_b200.so has ABI version {L.prk_abi_version()}, this package needs {ABI_VERSION}: '
                          'rebuild it with `python -m poserisk_release_b200.build`')
    _lib = L
    return L


class PeerExchangeError(PoseRiskError):
    """The peer-memory exchange is unavailable (CUDA IPC refused, no peer access, a peer timed out)."""


def check(rc: int):
    if rc != PRK_OK:
        L = lib()
        cls = PeerExchangeError if rc == PRK_ERR_PEER else PoseRiskError
        raise cls(f'{L.prk_strerror(rc).decode()} ({L.prk_last_error_detail().decode()})')


def launch_count() -> int:
    return int(lib().prk_launch_count())


def addinfo_array(add_infos) -> np.ndarray:
    """dict, or sequence of dicts, in additional_information.json layout -> int32 (T, 16).

    Missing keys raise KeyError exactly where the reference would (reba.py:59 etc.)."""
    if isinstance(add_infos, dict):
        add_infos = [add_infos]
    out = np.zeros((len(add_infos), 16), np.int32)
    for t, ai in enumerate(add_infos):
        for k, key in enumerate(REBA_KEYS):
            out[t, k] = _as_int(ai["REBA"][key], key)
        for k, key in enumerate(RULA_KEYS):
            out[t, 7 + k] = _as_int(ai["RULA"][key], key)
    return out


def _as_int(v, key):
    # the reference accumulates these into int64 numpy arrays (reba.py:123-129): a float
    # modifier raises a casting error there, so refuse it here too
    if isinstance(v, (bool, int, np.integer)):
        return int(v)
    raise TypeError(f'additional information "{key}" must be an integer, got {type(v).__name__}')
