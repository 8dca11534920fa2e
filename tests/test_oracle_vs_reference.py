"""Pin the C oracle against the LIVE unmodified reference (build container only;
skipped wherever /root/reference is absent, e.g. the GPU box)."""
import json
import os
import types

import numpy as np
import pytest

from ref_harness import reference_available, load_reference, ref_log_to_parts
from oracle import oracle
from poserisk_release_b200.model_provider import synthetic_smpl

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference tree not present')

USED = [3, 4, 5, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def _fuzz(rng, n):
    from golden.make_golden import fuzz_euler, expand_used
    return expand_used(fuzz_euler(rng, n))


def test_scores_big_fuzz_bit_exact():
    from golden.make_golden import random_info
    ref = load_reference()
    reba, rula = ref.REBA(), ref.RULA()
    rng = np.random.default_rng(77)
    n, blk = 20000, 500
    full = _fuzz(rng, n)
    dummy = np.zeros((blk, 1))
    for b in range(n // blk):
        ai = random_info(rng, wild=(b % 4 == 3))
        e = full[b * blk:(b + 1) * blk]
        rec = oracle.score_euler(e, ai)
        s, p = ref_log_to_parts(reba(e, dummy, ai), 'REBA')
        assert (rec['reba_score'] == s).all() and (rec['reba_parts'] == p).all()
        s, p = ref_log_to_parts(rula(e, dummy, ai), 'RULA')
        assert (rec['rula_score'] == s).all() and (rec['rula_parts'] == p).all()


def test_pose_to_score_pipeline_matches_reference_literal_path():
    """base.py:225-229,151,168 on 200 random f32 poses with the example JSON."""
    ref = load_reference()
    with open(os.path.join(ref.root, 'example', 'additional_information.json')) as f:
        ai = json.load(f)
    rng = np.random.default_rng(5)
    pose = rng.normal(0, 0.6, (200, 24, 3)).astype(np.float32)
    e_ref = np.stack([ref.coord_utils.axis_angle_to_euler_angle(p) for p in pose])
    rec, e = oracle.score_pose(pose, ai, want_euler=True)
    assert np.abs(e - e_ref).max() < 1e-9
    dummy = np.zeros((200, 1))
    s, p = ref_log_to_parts(ref.REBA()(e_ref, dummy, ai), 'REBA')
    assert (rec['reba_score'] == s).all() and (rec['reba_parts'] == p).all()
    s, p = ref_log_to_parts(ref.RULA()(e_ref, dummy, ai), 'RULA')
    assert (rec['rula_score'] == s).all() and (rec['rula_parts'] == p).all()


@pytest.mark.parametrize('gender', ['neutral', 'female', 'male'])
def test_smpl_forward_random(gender):
    import torch
    ref = load_reference()
    layer = ref.SMPL_Layer(gender=gender, model_root='unused')
    g = torch.Generator().manual_seed(11)
    pose = torch.randn(16, 72, generator=g) * 0.6
    betas = torch.randn(16, 10, generator=g)
    trans = torch.randn(16, 3, generator=g)
    v_ref, j_ref = layer(pose, betas, trans)
    v, j = oracle.smpl_forward(synthetic_smpl(gender), pose.numpy(), betas.numpy(), trans.numpy())
    assert np.abs(v - v_ref.numpy()).max() / np.abs(v_ref.numpy()).max() < 1e-5
    assert np.abs(j - j_ref.numpy()).max() / np.abs(j_ref.numpy()).max() < 1e-5


def test_euler_random_and_rotmat_roundtrip():
    ref = load_reference()
    rng = np.random.default_rng(3)
    p = rng.normal(0, 1.5, (500, 3)).astype(np.float32)
    e_ref = ref.coord_utils.axis_angle_to_euler_angle(p)
    e, bad = oracle.euler(p)
    assert not bad.any() and np.abs(e - e_ref).max() < 1e-9
