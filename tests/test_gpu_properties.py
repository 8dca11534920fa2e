"""Size-independent properties of the CUDA path at BASELINE.json's full batch sizes, where the
CPU oracle would take minutes: known answers for the identity pose, translation equivariance,
independence of a frame from its position in the batch (tile, lane, CTA), determinism, and the
joints-only path on 1M frames with the debug Euler sequences (configs 3 and 5)."""
import numpy as np
import pytest
import torch

from oracle import oracle
from poserisk_release_b200.model_provider import synthetic_smpl

pytestmark = pytest.mark.gpu

TOL = 1e-5
EXAMPLE_INFO = {"REBA": {"Legs_bilateral_weight_bearing/walking": 1, "Sitting": 1, "Load/Force Score": 0,
                         "Arm_supported_leaning_L": 0, "Arm_supported_leaning_R": 0, "Coupling": 0,
                         "Activity_Score": 0},
                "RULA": {"Arm_supported_leaning_L": 0, "Arm_supported_leaning_R": 0, "A_Muscle_use_L": 0,
                         "A_Muscle_use_R": 0, "A_Load/Force_L": 0, "A_Load/Force_R": 0,
                         "Legs_bilateral_weight_bearing": 0, "B_Muscle_use": 0, "B_Load/Force": 0}}


def relerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope='module')
def engine():
    from poserisk_release_b200 import PoseRiskEngine
    return PoseRiskEngine('cuda:0')


def test_identity_pose_known_answer(engine):
    """smpl_layer.py:93-95,145: with the identity pose the mesh is v_template + shapedirs.beta and the
    joints are J_regressor @ v_shaped; all Euler angles are 0, so the scores are those of the zero pose."""
    m = synthetic_smpl('neutral')
    B = 4096
    g = torch.Generator().manual_seed(11)
    betas = torch.randn(B, 10, generator=g)
    out = engine.run(torch.zeros(B, 72).cuda(), betas.cuda(), None, add_info=EXAMPLE_INFO)
    v_shaped = (m.v_template.reshape(1, -1).astype(np.float64) +
                betas.numpy().astype(np.float64) @ m.shapedirs.reshape(-1, 10).astype(np.float64).T).reshape(B, 6890, 3)
    j_ref = np.einsum('jv,bvc->bjc', np.asarray(m.J_regressor, np.float64), v_shaped)
    assert relerr(out['verts'].cpu().numpy(), v_shaped) < TOL
    assert relerr(out['joints'].cpu().numpy(), j_ref) < TOL
    rec = out['scores'].cpu().numpy()
    ref = oracle.score_pose(np.zeros((1, 72), np.float32), EXAMPLE_INFO)
    assert (rec == rec[0]).all()
    from poserisk_release_b200 import _runtime
    r0 = _runtime.records_to_numpy(out['scores'][:1])
    assert r0['reba_score'][0] == ref['reba_score'][0] and r0['rula_score'][0] == ref['rula_score'][0]


def test_translation_equivariance_and_score_invariance_full_batch(engine):
    """smpl_layer.py:153-155: a non-zero trans batch adds trans to every vertex and joint; the scores
    only see the pose."""
    B = 8192
    g = torch.Generator().manual_seed(12)
    pose = (torch.randn(B, 72, generator=g) * 0.35).cuda()
    betas = torch.randn(B, 10, generator=g).cuda()
    trans = (torch.randn(B, 3, generator=g) * 0.1).cuda()
    a = engine.run(pose, betas, None, add_info=EXAMPLE_INFO)
    b = engine.run(pose, betas, trans, add_info=EXAMPLE_INFO)
    dv = (b['verts'] - a['verts'] - trans[:, None, :]).abs().max().item()
    dj = (b['joints'] - a['joints'] - trans[:, None, :]).abs().max().item()
    scale = a['verts'].abs().max().item()
    assert dv < 4e-7 * scale + 1e-7 and dj < 4e-7 * scale + 1e-7      # one fp32 rounding of x + t
    assert torch.equal(a['scores'], b['scores'])


def test_frame_results_do_not_depend_on_batch_position(engine):
    """Frames are independent: the same frame gives bit-identical vertices, joints and
    scores wherever it sits in a batch (TMEM lane, frame tile, CTA) and whatever the batch size; two
    runs of the same batch are bit-identical (no atomics, fixed summation order)."""
    B = 4096
    g = torch.Generator().manual_seed(13)
    pose = (torch.randn(B, 72, generator=g) * 0.35).cuda()
    betas = torch.randn(B, 10, generator=g).cuda()
    trans = (torch.randn(B, 3, generator=g) * 0.1).cuda()
    full = engine.run(pose, betas, trans, add_info=EXAMPLE_INFO)
    again = engine.run(pose, betas, trans, add_info=EXAMPLE_INFO)
    for k in ('verts', 'joints', 'scores'):
        assert torch.equal(full[k], again[k])
    perm = torch.randperm(B, generator=g).cuda()
    shuf = engine.run(pose[perm], betas[perm], trans[perm], add_info=EXAMPLE_INFO)
    for k in ('verts', 'joints', 'scores'):
        assert torch.equal(shuf[k], full[k][perm])
    for lo, hi in ((0, 1), (100, 229), (3000, 4096)):
        part = engine.run(pose[lo:hi], betas[lo:hi], trans[lo:hi], add_info=EXAMPLE_INFO)
        for k in ('verts', 'joints', 'scores'):
            assert torch.equal(part[k], full[k][lo:hi])


def test_joints_only_one_million_frames_with_debug_euler(engine):
    """Configs 3/5 at full size on one GPU: 1M frames, joints-only, debug Euler sequences of four joints.
    Checked against the oracle on a random sample and through the thread-per-frame / warp-per-frame
    kernel equivalence on slices."""
    from poserisk_release_b200 import _runtime
    n = 1_000_000
    g = torch.Generator().manual_seed(14)
    pose = (torch.randn(n, 72, generator=g) * 0.35).cuda()
    betas = torch.randn(n, 10, generator=g).cuda()
    out = engine.run(pose, betas, None, add_info=EXAMPLE_INFO, want_verts=False)
    assert out['verts'] is None and tuple(out['joints'].shape) == (n, 24, 3)
    ids = [12, 16, 17, 3]                                  # Neck, L_Shoulder, R_Shoulder, Torso
    scores2, eul = engine.euler_debug(pose, ids, EXAMPLE_INFO)
    assert torch.equal(scores2, out['scores']) and tuple(eul.shape) == (n, 4, 3)
    # a sample against the oracle (scores bit-exact, Euler within 1e-3 degrees, joints within 1e-5)
    idx = torch.randint(0, n, (2000,), generator=g)
    p_s, b_s = pose[idx.cuda()].cpu().numpy(), betas[idx.cuda()].cpu().numpy()
    ref = oracle.score_pose(p_s, EXAMPLE_INFO)
    rec = _runtime.records_to_numpy(out['scores'][idx.cuda()])
    assert (rec['reba_score'] == ref['reba_score']).all() and (rec['rula_score'] == ref['rula_score']).all()
    assert (rec['reba_parts'] == ref['reba_parts']).all() and (rec['rula_parts'] == ref['rula_parts']).all()
    e_ref = oracle.euler(np.ascontiguousarray(p_s.reshape(-1, 24, 3)[:, ids]))[0]
    assert np.abs(eul[idx.cuda()].cpu().numpy() - e_ref).max() < 1e-3
    _, j_ref = oracle.smpl_forward(synthetic_smpl('neutral'), p_s, b_s, None, want_verts=False)
    assert relerr(out['joints'][idx.cuda()].cpu().numpy(), j_ref) < TOL
    # slices of the big batch (thread-per-frame kernel) against small launches (warp-per-frame kernel)
    for lo in (0, 499_000, n - 700):
        small = engine.run(pose[lo:lo + 700], betas[lo:lo + 700], None, add_info=EXAMPLE_INFO, want_verts=False)
        assert torch.equal(small['joints'], out['joints'][lo:lo + 700])
        assert torch.equal(small['scores'], out['scores'][lo:lo + 700])


@pytest.mark.parametrize('B', [1, 33, 129, 4096 + 77])
def test_vertex_stores_stay_inside_the_output(engine, B):
    """The vertex kernel stores frame rows in 8-byte pieces, part of them deferred into the next work unit
    (prk_fused.cu): a guard band on either side of the output must stay untouched, every output element
    must be written, and a second run into a NaN-filled buffer must give the same bits (no stale or missing
    store), for ragged batches whose last frame tile and last vertex tile are partial."""
    g = torch.Generator().manual_seed(B)
    pose = (torch.randn(B, 72, generator=g) * 0.35).cuda()
    betas = torch.randn(B, 10, generator=g).cuda()
    trans = (torch.randn(B, 3, generator=g) * 0.1).cuda()
    n, pad = B * 6890 * 3, 4096
    runs = []
    for fill in (float('nan'), 12345.0):
        flat = torch.full((n + 2 * pad,), fill, dtype=torch.float32, device='cuda')
        view = flat[pad:pad + n].view(B, 6890, 3)
        out = engine.run(pose, betas, trans, add_info=EXAMPLE_INFO, verts_out=view)
        torch.cuda.synchronize()
        assert out['verts'].data_ptr() == view.data_ptr()
        guard = torch.cat([flat[:pad], flat[pad + n:]])
        assert bool(torch.isnan(guard).all()) if fill != fill else bool((guard == fill).all())
        body = flat[pad:pad + n]
        assert not bool(torch.isnan(body).any())
        assert not bool((body == 12345.0).any())
        runs.append(body.clone())
    assert torch.equal(runs[0], runs[1])
