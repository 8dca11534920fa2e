"""Pin the C oracle against the LIVE unmodified reference (build container only;
skipped wherever /root/reference is absent, e.g. the GPU box)."""
import json
import os
import types

import numpy as np
import torch
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


def test_report_writers_match_reference_text(tmp_path):
    """report.py against the reference's own writers (lib/core/base.py:329-397, :158-165, :242-271 and
    lib/utils/vis_utils.py:9-16), run from their source: base.py / vis_utils.py cannot be imported here
    (matplotlib, SPIN, tracker), so the function bodies are compiled from the files as they are."""
    import ast, csv, filecmp, os.path as osp
    from ref_harness import REF_ROOT
    from poserisk_release_b200 import report

    def functions_of(path, names, cls=None):
        tree = ast.parse(open(path).read())
        body = tree.body
        if cls:
            body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
        picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
        mod = ast.Module(body=picked, type_ignores=[])
        from scipy.stats import mode

        class _Plt:                                # the score plot is out of scope: swallow it
            def __getattr__(self, _):
                return lambda *a, **k: None
        ns = {'np': np, 'csv': csv, 'osp': osp, 'mode': mode, 'plt': _Plt()}
        exec(compile(mod, path, 'exec'), ns)
        return ns

    base = functions_of(osp.join(REF_ROOT, 'lib', 'core', 'base.py'),
                        {'save_csv_pose_log', 'save_csv', 'post_processing'}, cls='Predictor')
    vis = functions_of(osp.join(REF_ROOT, 'lib', 'utils', 'vis_utils.py'), {'pose_to_str'})
    ref = load_reference()
    rng = np.random.default_rng(5)
    n = 40
    frames = np.sort(rng.choice(np.arange(3, 60), n, replace=False))
    timestamp = (0, frames, 64)
    euler = rng.uniform(-180, 180, (n, 24, 3))
    with open(os.path.join(REF_ROOT, 'example', 'additional_information.json')) as f:
        info = json.load(f)
    pose_str_ref = vis['pose_to_str'](euler)
    assert report.pose_to_str(euler) == pose_str_ref
    joints_upper = [j.upper() for j in ref.REBA().joint_name] if hasattr(ref.REBA(), 'joint_name') else None
    assert joints_upper is not None
    self_ = types.SimpleNamespace(debug_joints=['Neck', 'L_Shoulder', 'Torso'],
                                  smpl_model=types.SimpleNamespace(joints_name_upper=joints_upper))
    d_ref, d_new = tmp_path / 'ref', tmp_path / 'new'
    d_ref.mkdir(); d_new.mkdir()
    base['save_csv_pose_log'](self_, pose_str_ref, timestamp, str(d_ref))
    report.save_csv_pose_log(pose_str_ref, timestamp, str(d_new), self_.debug_joints, joints_upper)
    assert filecmp.cmp(d_ref / 'pose_log.csv', d_new / 'pose_log.csv', shallow=False)
    for title, scorer in (('REBA', ref.REBA(True)), ('RULA', ref.RULA(True))):
        results = scorer(euler, np.zeros((n, 1)), info)
        fin_ref, sc_ref, lg_ref = base['post_processing'](self_, results, scorer.eval_items, timestamp, str(d_ref), title=title)
        fin_new, sc_new, lg_new = report.post_processing(results)
        assert fin_new == fin_ref and (sc_new == sc_ref).all() and (lg_new == lg_ref).all()
        base['save_csv'](self_, pose_str_ref, timestamp, sc_ref, scorer.eval_items, lg_ref, scorer.log, str(d_ref), title=title)
        report.save_csv(pose_str_ref, timestamp, sc_new, scorer.eval_items, lg_new, scorer.log, str(d_new), title=title)
        for name in (title + '_score_log.csv', title + '_eval_pose_log.csv'):
            assert filecmp.cmp(d_ref / name, d_new / name, shallow=False), name
    # result text: the reference builds it inline (base.py:162-163 / :179-180); evaluate those very lines
    lines = open(osp.join(REF_ROOT, 'lib', 'core', 'base.py')).read().split('\n')
    for title, var in (('REBA', 'reba'), ('RULA', 'rula')):
        i = next(k for k, l in enumerate(lines) if 'data = f"AVG Score' in l and f'final_score_{var}' in l)
        expr = (lines[i] + '\n' + lines[i + 1]).strip()[len('data = '):]
        final = (4.123, 5.5, 7.0, 9, 4)
        ns = {f'final_score_{var}': final, f'{var}_action_level': 3, f'{var}_action_name': 'Medium risk.'}
        assert report.result_text(final, 3, 'Medium risk.', title) == eval(expr, ns)


def test_save_obj_matches_reference_bytes(tmp_path):
    """report.save_obj against the reference's own function (lib/utils/vis_utils.py:238-245), compiled from its
    source (vis_utils imports matplotlib/cv2 plotting helpers that are out of scope): float32 millimetre vertices as
    base.py:279-281 passes them, with and without faces."""
    import ast, filecmp, os.path as osp
    from ref_harness import REF_ROOT
    from poserisk_release_b200 import report
    path = osp.join(REF_ROOT, 'lib', 'utils', 'vis_utils.py')
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'save_obj']
    ns = {}
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, 'exec'), ns)
    m = synthetic_smpl('neutral')
    rng = np.random.default_rng(9)
    v = (rng.normal(0, 0.4, (6890, 3)).astype(np.float32) * 1000)
    v[0] = [0.0, -0.0, 1e-5]
    v[1] = [123456.789, 1e9, -3.4e-12]
    faces = m.faces.astype(np.uint32)
    ns['save_obj'](v, faces, str(tmp_path / 'ref.obj'))
    report.save_obj(v, faces, str(tmp_path / 'new.obj'))
    assert filecmp.cmp(tmp_path / 'ref.obj', tmp_path / 'new.obj', shallow=False)
    ns['save_obj'](v[:50], None, str(tmp_path / 'ref2.obj'))
    report.save_obj(torch.from_numpy(v[:50]), None, str(tmp_path / 'new2.obj'))
    assert filecmp.cmp(tmp_path / 'ref2.obj', tmp_path / 'new2.obj', shallow=False)
    # the int64 faces th_faces.numpy() gives (smpl.py:11) print the same text
    ns['save_obj'](v[:50], m.faces[:20].astype(np.int64), str(tmp_path / 'ref3.obj'))
    report.save_obj(v[:50], torch.from_numpy(m.faces[:20].astype(np.int64)), str(tmp_path / 'new3.obj'))
    assert filecmp.cmp(tmp_path / 'ref3.obj', tmp_path / 'new3.obj', shallow=False)


def test_smpl_wrapper_attributes_match_the_live_reference():
    """SMPL() of lib/utils/smpl.py:7-45 against ours: every public attribute the reference sets, same values and
    types (the three layers are compared through their buffers)."""
    import importlib, sys
    ref = load_reference()                       # puts lib/utils + lib/smplpytorch on sys.path, stubs ready_arguments
    ref_smpl = importlib.import_module('smpl')   # the reference's lib/utils/smpl.py
    assert ref_smpl.__file__.startswith(ref.root)
    r = ref_smpl.SMPL()
    from poserisk_release_b200.smpl import SMPL
    o = SMPL()
    public = [k for k in vars(r) if not k.startswith('_')]
    assert sorted(public) == sorted(k for k in vars(o) if not k.startswith('_'))
    for k in public:
        a, b = getattr(r, k), getattr(o, k)
        if k == 'layer':
            assert sorted(a) == sorted(b)
            for g in a:
                for name in ('th_betas', 'th_shapedirs', 'th_posedirs', 'th_v_template', 'th_J_regressor', 'th_weights', 'th_faces'):
                    ta, tb = getattr(a[g], name), getattr(b[g], name)
                    assert ta.dtype == tb.dtype and ta.shape == tb.shape and torch.equal(ta, tb), (g, name)
                assert a[g].kintree_parents == b[g].kintree_parents and a[g].num_joints == b[g].num_joints
                assert a[g].model_path == b[g].model_path and a[g].gender == b[g].gender and a[g].center_idx == b[g].center_idx
                assert torch.equal(a[g].vertice_segmentation, b[g].vertice_segmentation)
        elif isinstance(a, np.ndarray):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), k
        else:
            assert type(a) is type(b) and a == b, k
