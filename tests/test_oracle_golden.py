"""Pin the C oracle (oracle/poserisk_oracle.c) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  Runs on every box, no GPU."""
import os

import numpy as np
import pytest

from oracle import oracle
from poserisk_release_b200.model_provider import synthetic_smpl

USED = [3, 4, 5, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def relerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_euler_cfg1_f32_pose(golden):
    g = golden['euler']
    e, bad = oracle.euler(g['cfg1_pose'])
    assert not bad.any()
    # same libm-level closed form: agreement is ~1 ulp of a degree value
    assert np.abs(e - g['cfg1_euler']).max() < 1e-9


def test_euler_f64_pose_and_specials(golden):
    g = golden['euler']
    e, _ = oracle.euler(g['f64_pose'])
    assert np.abs(e - g['f64_euler']).max() < 1e-9
    e, _ = oracle.euler(g['special64'])
    assert np.abs(e - g['special64_euler']).max() < 1e-9
    e, _ = oracle.euler(g['special32'])
    assert np.abs(e - g['special32_euler']).max() < 1e-9
    # known answers
    e, _ = oracle.euler(np.array([[3.14, 0, 0], [0, np.pi / 2, 0]]))
    assert abs(e[0, 0] - 179.9087) < 1e-4 and abs(e[1, 1] - 90.0) < 1e-6


def test_scores_cfg1(golden):
    e, s = golden['euler'], golden['scores']
    rec = oracle.score_euler(e['cfg1_euler'], _info_dicts(s['example_info'][None])[0])
    assert (rec['reba_score'] == s['cfg1_reba_score']).all()
    assert (rec['rula_score'] == s['cfg1_rula_score']).all()
    assert (rec['reba_parts'] == s['cfg1_reba_parts']).all()
    assert (rec['rula_parts'] == s['cfg1_rula_parts']).all()
    # fused pose->score entry must agree with the two-step path
    rec2 = oracle.score_pose(e['cfg1_pose'], _info_dicts(s['example_info'][None])[0])
    for k in ('reba_score', 'rula_score', 'reba_parts', 'rula_parts'):
        assert (rec2[k] == rec[k]).all()


def _info_dicts(rows):
    out = []
    for r in rows:
        out.append({'REBA': dict(zip(oracle.REBA_KEYS, (int(x) for x in r[:7]))),
                    'RULA': dict(zip(oracle.RULA_KEYS, (int(x) for x in r[7:])))})
    return out


def test_scores_fuzz_thresholds_nan(golden):
    s = golden['scores']
    e12 = s['fuzz_euler12']
    full = np.zeros((e12.shape[0], 24, 3))
    full[:, USED] = e12
    rec = oracle.score_euler(full, _info_dicts(s['fuzz_info']), s['fuzz_track'])
    assert (rec['reba_score'] == s['fuzz_reba_score']).all()
    assert (rec['rula_score'] == s['fuzz_rula_score']).all()
    assert (rec['reba_parts'] == s['fuzz_reba_parts']).all()
    assert (rec['rula_parts'] == s['fuzz_rula_parts']).all()


@pytest.mark.parametrize('case', ['A', 'B', 'C', 'D', 'E'])
def test_smpl_forward(golden, case):
    g = golden['smpl_forward']
    tol = 1e-5  # BASELINE.json: vertices and joints within 1e-5 relative (max-abs / max-mag)
    if case == 'A':
        v, j = oracle.smpl_forward(synthetic_smpl('neutral'), g['A_pose'], g['A_betas'], g['A_trans'])
        assert relerr(v, g['A_verts']) < tol and relerr(j, g['A_joints']) < tol
    elif case == 'B':
        v, j = oracle.smpl_forward(synthetic_smpl('female'), g['B_pose'])
        assert relerr(v[:, ::5], g['B_verts5']) < tol and relerr(j, g['B_joints']) < tol
    elif case == 'C':
        v, j = oracle.smpl_forward(synthetic_smpl('male'), g['C_pose'], g['C_betas'],
                                   np.zeros((3, 3), np.float32), center_idx=0)
        assert relerr(v[:, ::5], g['C_verts5']) < tol and relerr(j, g['C_joints']) < tol
        assert np.abs(j[:, 0]).max() == 0.0
    elif case == 'D':
        m = synthetic_smpl('neutral')
        v, j = oracle.smpl_forward(m, np.zeros((1, 72), np.float32), g['D_betas'])
        assert relerr(v[:, ::5], g['D_verts5']) < tol and relerr(j, g['D_joints']) < tol
        # known answer: identity pose => verts = v_template + shapedirs.beta (smpl_layer.py:93-95)
        ka = m.v_template + m.shapedirs @ g['D_betas'][0]
        assert relerr(v[0], ka) < 1e-6
    else:
        v, j = oracle.smpl_forward(synthetic_smpl('neutral'), g['E_pose'], g['E_betas'])
        assert relerr(v[:, ::5], g['E_verts5']) < tol and relerr(j, g['E_joints']) < tol


def test_get_joint_cam_golden(golden):
    g = golden['smpl_forward']
    pose = g['F_pose'].copy()
    pose[:, 0] = np.array([3.14, 0, 0], np.float32)        # coord_utils.py:10,13
    assert np.array_equal(pose, g['F_pose_after'])
    _, j = oracle.smpl_forward(synthetic_smpl('neutral'), pose.reshape(-1, 72), want_verts=False)
    jc = j * 1000
    jc = jc - jc[:, :1]
    assert np.abs(jc - g['F_joint_cam']).max() / np.abs(g['F_joint_cam']).max() < 1e-5


def test_rot_to_angle_golden():
    """oracle rot_to_angle against the reference's rot_to_angle (cv2.Rodrigues 4.13.0) on float32 and
    float64 matrices incl. half turns, near-identity and not-quite-orthonormal input; then the chain
    rot_to_angle -> axis_angle_to_euler_angle of base.py:225-229."""
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'rotmat.npz'))
    out, bad = oracle.rot_to_angle(g['rotmat'])
    assert out.dtype == np.float32 and not bad.any()
    assert np.abs(out - g['pose']).max() <= 2.4e-7                       # one float32 ulp at |rvec| < pi
    assert (out.view(np.uint32) == g['pose'].view(np.uint32)).mean() > 0.998
    o64, _ = oracle.rot_to_angle(g['rotmat64'])
    assert np.abs(o64 - g['pose64']).max() < 1e-6
    e, _ = oracle.euler(out)
    assert np.abs(e - g['euler']).max() < 1e-3
    # singular input is flagged, not propagated
    _, bad = oracle.rot_to_angle(np.zeros((1, 3, 3), np.float32))
    assert bad.all()
