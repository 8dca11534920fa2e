"""Generate tests/golden/*.npz by EXECUTING the unmodified reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference publishes no golden vectors; these files are the
pin for the C oracle (oracle/poserisk_oracle.c) and, through it, the CUDA path.
Reference entry points exercised:
  SMPL_Layer.forward            lib/smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158
  get_joint_cam                 lib/utils/coord_utils.py:7-21
  axis_angle_to_euler_angle     lib/utils/coord_utils.py:83-95
  REBA.__call__ / RULA.__call__ lib/utils/reba.py:50-81, lib/utils/rula.py:66-98
with the synthetic SMPL-shaped model of poserisk_release_b200.model_provider.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from ref_harness import load_reference, ref_log_to_parts  # noqa: E402

THRESHOLDS = np.array([0, 1, 5, 10, 15, 20, 30, 45, 60, 70, 90, 100, 110], np.float64)
USED_JOINTS = [3, 4, 5, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def fuzz_euler(rng, n):
    """Euler degrees uniform in +-180 with ~35% of used entries snapped onto / next to a threshold."""
    e = rng.uniform(-180.0, 180.0, (n, 12, 3))
    snap = rng.random((n, 12, 3)) < 0.35
    thr = rng.choice(THRESHOLDS, (n, 12, 3)) * rng.choice([-1.0, 1.0], (n, 12, 3))
    eps = rng.choice([0.0, 1e-9, -1e-9, 1e-4, -1e-4], (n, 12, 3))
    e = np.where(snap, thr + eps, e)
    return e


def expand_used(e12):
    full = np.zeros((e12.shape[0], 24, 3), np.float64)
    full[:, USED_JOINTS] = e12
    return full


def random_info(rng, wild=False):
    lo, hi = (-1, 5) if wild else (0, 4)
    r = lambda a, b: int(rng.integers(a, b))
    return {
        "REBA": {"Legs_bilateral_weight_bearing/walking": r(0 if wild else 1, 4 if wild else 3),
                 "Sitting": r(0, 2), "Load/Force Score": r(lo, hi),
                 "Arm_supported_leaning_L": r(0, 2), "Arm_supported_leaning_R": r(0, 2),
                 "Coupling": r(lo, hi), "Activity_Score": r(lo, hi)},
        "RULA": {"Arm_supported_leaning_L": r(0, 2), "Arm_supported_leaning_R": r(0, 2),
                 "A_Muscle_use_L": r(0, 2), "A_Muscle_use_R": r(0, 2),
                 "A_Load/Force_L": r(lo, hi), "A_Load/Force_R": r(lo, hi),
                 "Legs_bilateral_weight_bearing": r(0 if wild else 1, 4 if wild else 3),
                 "B_Muscle_use": r(0, 2), "B_Load/Force": r(lo, hi)}}


def info_to_row(ai):
    from oracle.oracle import addinfo_array
    return addinfo_array(ai)[0]


def main():
    ref = load_reference()
    with open(os.path.join(ref.root, 'example', 'additional_information.json')) as f:
        example_info = json.load(f)
    out = {}

    # ---------------- SMPL forward ----------------
    smpl = {}
    layers = {g: ref.SMPL_Layer(gender=g, model_root='unused') for g in ('neutral', 'female', 'male')}
    g = torch.Generator().manual_seed(0)
    # A: everything on, neutral
    pose = torch.randn(4, 72, generator=g) * 0.35
    betas = torch.randn(4, 10, generator=g)
    trans = torch.randn(4, 3, generator=g) * 0.1
    v, j = layers['neutral'](pose, betas, trans)
    smpl.update(A_pose=pose.numpy(), A_betas=betas.numpy(), A_trans=trans.numpy(), A_verts=v.numpy(), A_joints=j.numpy())
    # B: defaults (zero betas, no trans), female, every 5th vertex kept
    pose = torch.randn(3, 72, generator=g) * 0.5
    v, j = layers['female'](pose)
    smpl.update(B_pose=pose.numpy(), B_verts5=v.numpy()[:, ::5], B_joints=j.numpy())
    # C: betas, zero trans, center_idx=0, male
    lay = ref.SMPL_Layer(center_idx=0, gender='male', model_root='unused')
    pose = torch.randn(3, 72, generator=g) * 0.35
    betas = torch.randn(3, 10, generator=g) * 2.0
    v, j = lay(pose, betas, torch.zeros(3, 3))
    smpl.update(C_pose=pose.numpy(), C_betas=betas.numpy(), C_verts5=v.numpy()[:, ::5], C_joints=j.numpy())
    # D: identity pose with betas (known answer: v_template + shapedirs.beta)
    pose = torch.zeros(1, 72)
    betas = torch.randn(1, 10, generator=g)
    v, j = layers['neutral'](pose, betas)
    smpl.update(D_betas=betas.numpy(), D_verts5=v.numpy()[:, ::5], D_joints=j.numpy())
    # E: large rotations (|aa| up to ~2*pi) and an exact-zero joint
    pose = (torch.rand(3, 72, generator=g) - 0.5) * 7.0
    pose[:, 9:12] = 0.0
    betas = torch.randn(3, 10, generator=g)
    v, j = layers['neutral'](pose, betas)
    smpl.update(E_pose=pose.numpy(), E_betas=betas.numpy(), E_verts5=v.numpy()[:, ::5], E_joints=j.numpy())
    # F: get_joint_cam (coord_utils.py:7-21): mutates input, root=[3.14,0,0], mm, pelvis-relative
    rng = np.random.default_rng(1234)
    poses = rng.normal(0.0, 0.35, (300, 24, 3)).astype(np.float32)
    cfg1_pose = poses.copy()
    smpl_model = types.SimpleNamespace(layer=layers)
    jc_in = poses[:24].copy()
    jc = ref.coord_utils.get_joint_cam(jc_in, smpl_model)
    smpl.update(F_pose=poses[:24].copy(), F_pose_after=jc_in, F_joint_cam=jc)
    np.savez_compressed(os.path.join(HERE, 'smpl_forward.npz'), **smpl)

    # ---------------- Euler ----------------
    eul = {}
    e32 = np.stack([ref.coord_utils.axis_angle_to_euler_angle(p) for p in cfg1_pose])
    eul.update(cfg1_pose=cfg1_pose, cfg1_euler=e32)
    p64 = rng.normal(0.0, 1.2, (64, 24, 3))
    e64 = np.stack([ref.coord_utils.axis_angle_to_euler_angle(p) for p in p64])
    eul.update(f64_pose=p64, f64_euler=e64)
    special = np.array([[0, 0, 0], [3.14, 0, 0], [0, np.pi / 2, 0], [0, -np.pi / 2, 0],
                        [1e-20, 0, 0], [1e-9, -1e-9, 1e-9], [np.pi, 0, 0], [0, 0, np.pi],
                        [4.0, 3.0, -5.0], [0.3, 1.5707963, 0.0], [2e-16, 0, 0], [1e-17, 1e-17, 0]],
                       np.float64)
    eul.update(special64=special, special64_euler=ref.coord_utils.axis_angle_to_euler_angle(special))
    sp32 = special.astype(np.float32)
    eul.update(special32=sp32, special32_euler=ref.coord_utils.axis_angle_to_euler_angle(sp32))
    np.savez_compressed(os.path.join(HERE, 'euler.npz'), **eul)

    # ---------------- scores ----------------
    sc = {}
    reba, rula = ref.REBA(), ref.RULA()
    dummy = np.zeros((300, 1))
    s, p = ref_log_to_parts(reba(e32, dummy, example_info), 'REBA')
    sc.update(cfg1_reba_score=s.astype(np.int16), cfg1_reba_parts=p.astype(np.uint8))
    s, p = ref_log_to_parts(rula(e32, dummy, example_info), 'RULA')
    sc.update(cfg1_rula_score=s.astype(np.int16), cfg1_rula_parts=p.astype(np.uint8))
    sc['example_info'] = info_to_row(example_info)

    n, blk = 3072, 128
    rng = np.random.default_rng(2024)
    e12 = fuzz_euler(rng, n)
    # a few non-finite entries: NaN/inf fall to the `else` arms
    for k in range(48):
        e12[rng.integers(n), rng.integers(12), rng.integers(3)] = [np.nan, np.inf, -np.inf][k % 3]
    full = expand_used(e12)
    infos, track = [], np.zeros(n, np.int32)
    rs, rp, us, up = [], [], [], []
    dummy = np.zeros((blk, 1))
    for b in range(n // blk):
        ai = random_info(rng, wild=(b % 3 == 2))
        infos.append(info_to_row(ai))
        track[b * blk:(b + 1) * blk] = b
        s, p = ref_log_to_parts(reba(full[b * blk:(b + 1) * blk], dummy, ai), 'REBA')
        rs.append(s); rp.append(p)
        s, p = ref_log_to_parts(rula(full[b * blk:(b + 1) * blk], dummy, ai), 'RULA')
        us.append(s); up.append(p)
    sc.update(fuzz_euler12=e12, fuzz_info=np.stack(infos), fuzz_track=track,
              fuzz_reba_score=np.concatenate(rs).astype(np.int16), fuzz_reba_parts=np.concatenate(rp).astype(np.uint8),
              fuzz_rula_score=np.concatenate(us).astype(np.int16), fuzz_rula_parts=np.concatenate(up).astype(np.uint8))
    np.savez_compressed(os.path.join(HERE, 'scores.npz'), **sc)
    for f in ('smpl_forward.npz', 'euler.npz', 'scores.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
