"""Generate tests/golden/rotmat.npz by EXECUTING the unmodified reference's rot_to_angle
(lib/utils/coord_utils.py:24-30 -> cv2.Rodrigues, opencv-python 4.13.0 in the build container)
and, on its output, axis_angle_to_euler_angle (coord_utils.py:83-95), as lib/core/base.py:225-229
chains them.  Run in the build container only:  python tests/golden/make_golden_rotmat.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from ref_harness import load_reference  # noqa: E402


def main():
    ref = load_reference()
    rng = np.random.default_rng(77)
    frames = []
    for f in range(96):
        aa = rng.normal(0.0, 0.6, (24, 3))
        if f % 8 == 1:                                   # half turns and almost half turns
            aa = aa / np.linalg.norm(aa, axis=1, keepdims=True) * (np.pi - rng.choice([0.0, 1e-7, 1e-4, 1e-2], (24, 1)))
        if f % 8 == 2:                                   # almost no rotation
            aa = aa * 1e-6
        R = np.stack([cv2.Rodrigues(a)[0] for a in aa])
        if f % 8 == 3:                                   # not exactly orthonormal (what a network regresses)
            R = R + rng.normal(0.0, 1e-3, R.shape)
        if f % 8 == 4:
            R[0] = np.eye(3); R[1] = np.diag([1.0, -1.0, -1.0]); R[2] = np.diag([-1.0, 1.0, -1.0]); R[3] = np.diag([-1.0, -1.0, 1.0])
        frames.append(R.astype(np.float32))
    rotmat = np.stack(frames)                            # (96, 24, 3, 3) float32, like SPIN's pred_rotmat
    pose = np.stack([ref.coord_utils.rot_to_angle(r) for r in rotmat])
    assert pose.dtype == np.float32 and pose.shape == (96, 24, 3)
    euler = np.stack([ref.coord_utils.axis_angle_to_euler_angle(p) for p in pose])
    rot64 = rotmat[:8].astype(np.float64)
    pose64 = np.stack([ref.coord_utils.rot_to_angle(r) for r in rot64])
    np.savez_compressed(os.path.join(HERE, 'rotmat.npz'), rotmat=rotmat, pose=pose, euler=euler, rotmat64=rot64,
                        pose64=pose64, cv2_version=np.array(cv2.__version__))
    print('wrote rotmat.npz', rotmat.shape, 'cv2', cv2.__version__)


if __name__ == '__main__':
    main()
