"""ctypes binding of the CPU oracle (oracle/poserisk_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of poserisk_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'poserisk_oracle.c')
_SO = os.path.join(_HERE, 'libposerisk_oracle.so')

REBA_KEYS = ("Legs_bilateral_weight_bearing/walking", "Sitting", "Load/Force Score",
             "Arm_supported_leaning_L", "Arm_supported_leaning_R", "Coupling", "Activity_Score")
RULA_KEYS = ("Arm_supported_leaning_L", "Arm_supported_leaning_R", "A_Muscle_use_L",
             "A_Muscle_use_R", "A_Load/Force_L", "A_Load/Force_R",
             "Legs_bilateral_weight_bearing", "B_Muscle_use", "B_Load/Force")

REC_DTYPE = np.dtype([('reba_score', '<i2'), ('rula_score', '<i2'), ('reba_parts', 'u1', (9,)),
                      ('rula_parts', 'u1', (11,)), ('flags', 'u1'), ('pad', 'u1', (7,))])
assert REC_DTYPE.itemsize == 32


def build(force: bool = False) -> str:
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ['gcc', '-O3', '-march=x86-64-v3', '-fopenmp', '-fPIC', '-shared', '-std=c11', '-ffp-contract=off',
               '-o', _SO, _SRC, '-lm']
        subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_abi_version.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def addinfo_array(add_infos) -> np.ndarray:
    """dict or list of dicts (additional_information.json layout) -> int32 (T,16)."""
    if isinstance(add_infos, dict):
        add_infos = [add_infos]
    out = np.zeros((len(add_infos), 16), np.int32)
    for t, ai in enumerate(add_infos):
        out[t, :7] = [int(ai["REBA"][k]) for k in REBA_KEYS]
        out[t, 7:] = [int(ai["RULA"][k]) for k in RULA_KEYS]
    return out


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def euler(pose: np.ndarray):
    """axis_angle_to_euler_angle (coord_utils.py:83-95) over (...,3) axis-angles."""
    pose = np.ascontiguousarray(pose)
    assert pose.dtype in (np.float32, np.float64) and pose.shape[-1] == 3
    n = pose.size // 3
    out = np.empty(pose.shape, np.float64)
    bad = np.empty(n, np.uint8)
    lib().orc_euler(_p(pose), C.c_int(pose.dtype == np.float32), C.c_int64(n), _p(out), _p(bad))
    return out, bad.reshape(pose.shape[:-1])


def rot_to_angle(rotmat: np.ndarray):
    """rot_to_angle (coord_utils.py:24-30) over (...,3,3) rotation matrices; result has their dtype."""
    R = np.ascontiguousarray(rotmat)
    assert R.dtype in (np.float32, np.float64) and R.shape[-2:] == (3, 3)
    n = R.size // 9
    out = np.empty(R.shape[:-2] + (3,), R.dtype)
    bad = np.empty(n, np.uint8)
    lib().orc_rot_to_angle(_p(R), C.c_int(R.dtype == np.float32), C.c_int64(n), _p(out), _p(bad))
    return out, bad.reshape(R.shape[:-2])


def score_euler(euler_deg: np.ndarray, add_infos, track_of_frame=None) -> np.ndarray:
    e = np.ascontiguousarray(euler_deg, np.float64).reshape(-1, 24, 3)
    info = addinfo_array(add_infos)
    tr = None if track_of_frame is None else np.ascontiguousarray(track_of_frame, np.int32)
    out = np.zeros(e.shape[0], REC_DTYPE)
    lib().orc_score_euler(_p(e), _p(info), _p(tr), C.c_int64(e.shape[0]), _p(out))
    return out


def score_pose(pose: np.ndarray, add_infos, track_of_frame=None, want_euler=False):
    pose = np.ascontiguousarray(pose)
    assert pose.dtype in (np.float32, np.float64)
    p = pose.reshape(-1, 72)
    info = addinfo_array(add_infos)
    tr = None if track_of_frame is None else np.ascontiguousarray(track_of_frame, np.int32)
    out = np.zeros(p.shape[0], REC_DTYPE)
    eul = np.empty((p.shape[0], 24, 3), np.float64) if want_euler else None
    lib().orc_score_pose(_p(p), C.c_int(pose.dtype == np.float32), _p(info), _p(tr),
                         C.c_int64(p.shape[0]), _p(out), _p(eul))
    return (out, eul) if want_euler else out


def smpl_forward(model, pose, betas=None, trans=None, center_idx=None, want_verts=True):
    """SMPL_Layer.forward (smpl_layer.py:65-158) on a SMPLModelData; numpy float32 in/out."""
    f32 = np.float32
    pose = np.ascontiguousarray(pose, f32).reshape(-1, 72)
    B = pose.shape[0]
    betas = None if betas is None else np.ascontiguousarray(betas, f32).reshape(B, 10)
    trans = None if trans is None else np.ascontiguousarray(trans, f32).reshape(B, 3)
    vt = np.ascontiguousarray(model.v_template, f32)
    sd = np.ascontiguousarray(model.shapedirs, f32)
    pd = np.ascontiguousarray(model.posedirs, f32)
    jr = np.ascontiguousarray(model.J_regressor, f32)
    w = np.ascontiguousarray(model.weights, f32)
    parents = np.array([0] + [int(x) for x in model.parents[1:]], np.int32)
    mb = np.ascontiguousarray(model.betas, f32)
    verts = np.empty((B, 6890, 3), f32) if want_verts else None
    joints = np.empty((B, 24, 3), f32)
    lib().orc_smpl_forward(_p(vt), _p(sd), _p(pd), _p(jr), _p(w), _p(parents), _p(mb), _p(pose),
                           _p(betas), _p(trans), C.c_int(-1 if center_idx is None else center_idx),
                           C.c_int64(B), _p(verts), _p(joints))
    return verts, joints


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(n))
    return num_threads()
