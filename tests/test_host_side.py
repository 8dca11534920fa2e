"""CPU-only tests: C-ABI surface, host logic of the drop-ins, sharding over gloo."""
import ctypes
import json
import os
import pickle
import re
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/*.h declares."""
    from poserisk_release_b200.build import build_library
    lib = build_library()
    L = ctypes.CDLL(lib)
    hdr = open(os.path.join(ROOT, 'include', 'poserisk_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(prk_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(L, sym), sym
    from poserisk_release_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    L.prk_abi_version.restype = ctypes.c_int
    assert L.prk_abi_version() == _lib.ABI_VERSION == 2
    L.prk_strerror.restype = ctypes.c_char_p
    assert L.prk_strerror(0) == b'ok' and b'workspace' in L.prk_strerror(3)
    # struct sizes the Python side relies on
    assert _lib.REC_DTYPE.itemsize == 32


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import or load it."""
    pkg = os.path.join(ROOT, 'poserisk_release_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.lower(), os.path.join(dirpath, f)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from poserisk_release_b200 import SMPL_Layer, REBA
    lay = SMPL_Layer(gender='neutral', model_root='unused')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        lay(torch.zeros(1, 72))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        REBA()(np.zeros((1, 24, 3)), np.zeros((1, 1)), {"REBA": {}, "RULA": {}})


def test_smpl_layer_attributes_match_reference_surface():
    import torch
    from poserisk_release_b200 import SMPL_Layer
    lay = SMPL_Layer(center_idx=2, gender='female', model_root='some/root')
    assert lay.model_path == os.path.join('some/root', 'SMPL_FEMALE.pkl')
    assert lay.center_idx == 2 and lay.gender == 'female' and lay.num_joints == 24
    assert tuple(lay.th_betas.shape) == (1, 10) and tuple(lay.th_shapedirs.shape) == (6890, 3, 10)
    assert tuple(lay.th_posedirs.shape) == (6890, 3, 207) and tuple(lay.th_v_template.shape) == (1, 6890, 3)
    assert tuple(lay.th_J_regressor.shape) == (24, 6890) and tuple(lay.th_weights.shape) == (6890, 24)
    assert lay.th_faces.dtype == torch.int64 and lay.th_faces.shape[1] == 3
    assert lay.kintree_parents[1:] == [0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
    assert tuple(lay.vertice_segmentation.shape) == (6890,)
    assert lay.kintree_table.shape == (2, 24)


def test_pkl_loader_without_chumpy(tmp_path):
    """SMPL pickles hold chumpy objects; the loader must read them with chumpy absent."""
    import scipy.sparse as sp
    from poserisk_release_b200.model_provider import load_smpl_pkl, synthetic_smpl
    m = synthetic_smpl('neutral')
    ch_mod, ch_ch = types.ModuleType('chumpy'), types.ModuleType('chumpy.ch')

    class Ch:
        def __init__(self, x):
            self.x = np.asarray(x)

        def __getstate__(self):
            return {'x': self.x, 'dterms': ('x',)}
    Ch.__module__, Ch.__qualname__ = 'chumpy.ch', 'Ch'
    ch_ch.Ch = Ch
    ch_mod.ch = ch_ch
    sys.modules['chumpy'], sys.modules['chumpy.ch'] = ch_mod, ch_ch
    try:
        dd = {'v_template': m.v_template.astype(np.float64), 'shapedirs': Ch(m.shapedirs.astype(np.float64)),
              'posedirs': m.posedirs.astype(np.float64), 'weights': m.weights.astype(np.float64),
              'J_regressor': sp.csc_matrix(m.J_regressor.astype(np.float64)), 'f': m.faces.astype(np.uint32),
              'kintree_table': m.kintree_table, 'bs_type': 'lrotmin'}
        path = tmp_path / 'SMPL_NEUTRAL.pkl'
        with open(path, 'wb') as f:
            pickle.dump(dd, f, protocol=2)
    finally:
        del sys.modules['chumpy'], sys.modules['chumpy.ch']
    got = load_smpl_pkl(str(path))
    assert not got.synthetic
    for k in ('v_template', 'shapedirs', 'posedirs', 'weights', 'J_regressor'):
        assert np.array_equal(getattr(got, k), getattr(m, k)), k
    assert np.array_equal(got.betas, np.zeros(10, np.float32))
    assert got.parents[1:] == m.parents[1:]


def test_addinfo_validation():
    from poserisk_release_b200 import _lib
    from oracle import oracle
    ok = {"REBA": {k: 1 for k in _lib.REBA_KEYS}, "RULA": {k: 2 for k in _lib.RULA_KEYS}}
    a = _lib.addinfo_array([ok, ok])
    assert a.shape == (2, 16) and a.dtype == np.int32 and (a[:, :7] == 1).all() and (a[:, 7:] == 2).all()
    assert np.array_equal(a, oracle.addinfo_array([ok, ok]))
    bad = json.loads(json.dumps(ok))
    del bad["RULA"]["B_Muscle_use"]
    with pytest.raises(KeyError):
        _lib.addinfo_array(bad)
    bad = json.loads(json.dumps(ok))
    bad["REBA"]["Coupling"] = 1.5
    with pytest.raises(TypeError):
        _lib.addinfo_array(bad)


def test_aggregate_from_histogram_matches_post_processing():
    """base.py:260-271 on random score sequences, including n < 10 (top-10% mean is nan)."""
    from scipy.stats import mode
    from poserisk_release_b200.pipeline import aggregate_from_histogram
    rng = np.random.default_rng(0)
    for n in (1, 3, 9, 10, 11, 57, 1000, 4097):
        scores = rng.integers(1, 13, n)
        hist = np.bincount(scores + 16, minlength=64)
        s = np.sort(scores)[::-1]
        with np.errstate(all='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                expect = (round(s.mean(), 3), round(s[:n // 2].mean(), 3), round(s[:n // 10].mean(), 3),
                          round(s.max(), 3), mode(s).mode.item())
        got = aggregate_from_histogram(hist, n)
        for g, e in zip(got, expect):
            assert (np.isnan(g) and np.isnan(e)) or g == pytest.approx(e, abs=1e-12)


def test_debug_angle_logs_match_reference():
    """The debug `log` strings of REBA/RULA (reba.py:142-390, rula.py:198-418), including the
    `angle4=1` leak (rula.py:183,198) and the crossed variables in the abducted logs."""
    from ref_harness import reference_available, load_reference
    if not reference_available():
        pytest.skip('reference tree not present')
    from golden.make_golden import fuzz_euler, expand_used
    from poserisk_release_b200.reba import REBA
    from poserisk_release_b200.rula import RULA
    ref = load_reference()
    info = json.load(open(os.path.join(ref.root, 'example', 'additional_information.json')))
    e = expand_used(fuzz_euler(np.random.default_rng(1), 400))
    dummy = np.zeros((400, 1))
    r_reba, r_rula = ref.REBA(debug=True), ref.RULA(debug=True)
    r_reba(e, dummy, info)
    r_rula(e, dummy, info)
    mine_reba, mine_rula = REBA(debug=True), RULA(debug=True)
    for i in range(400):
        a, b = mine_reba._angle_log(e[i]), r_reba.log[i]
        assert a == b and list(a) == list(b)
        a, b = mine_rula._angle_log(e[i]), r_rula.log[i]
        assert a == b and list(a) == list(b)
    for s in (1, 2, 3, 5, 7, 8, 10, 11, 15, 4.4, 7.5):
        assert mine_reba.action_level(s) == r_reba.action_level(s)
        assert mine_rula.action_level(s) == r_rula.action_level(s)
    assert mine_reba.eval_items == r_reba.eval_items and mine_rula.eval_items == r_rula.eval_items
    assert np.array_equal(mine_reba.table_a, r_reba.table_a) and np.array_equal(mine_reba.table_c, r_reba.table_c)
    assert np.array_equal(mine_rula.table_a, r_rula.table_a) and np.array_equal(mine_rula.table_b, r_rula.table_b)


def test_shard_ranges_cover_everything():
    from poserisk_release_b200.distributed import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from poserisk_release_b200.distributed import shard_range, all_gather_rows, ScoreExchange, run_sharded, shard_tracks
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
rank, world = dist.get_rank(), dist.get_world_size()
for n in (10, 11, 4096):
    full = (torch.arange(n * 32, dtype=torch.int64) % 251).to(torch.uint8).reshape(n, 32)
    lo, hi = shard_range(n, rank, world)
    got = all_gather_rows(full[lo:hi].clone(), n)
    assert got.shape == full.shape and torch.equal(got, full), (n, rank)

# run_sharded with a stand-in engine: scores AND debug Euler sequences are gathered into frame order
class FakeEngine:
    device = None
    def run(self, pose, betas, trans, add_info=None, track_of_frame=None, want_verts=False, debug_joints=None,
            exchange=None, frame_offset=0):
        n = pose.shape[0]
        assert exchange is not None and exchange.native_handles(len(debug_joints)) == (None, None)   # CPU: collective
        scores = (pose[:, :32] * 7).to(torch.uint8)
        euler = pose[:, :len(debug_joints) * 3].double().reshape(n, len(debug_joints), 3) + 0.5
        return {'scores': scores, 'euler': euler, 'joints': pose[:, :72].reshape(n, 24, 3), 'verts': None}
n = 1001
pose = torch.arange(n * 72, dtype=torch.float32).reshape(n, 72) % 13
out = run_sharded(FakeEngine(), pose, None, None, add_info=None, debug_joints=[12, 16, 17, 3])
assert out['transport'] == 'nccl' and out['range'] == shard_range(n, rank, world)
assert torch.equal(out['scores'], (pose[:, :32] * 7).to(torch.uint8))
assert torch.equal(out['euler'], pose[:, :12].double().reshape(n, 4, 3) + 0.5)

# config 4: whole tracks per rank, ragged shards gathered with explicit sizes
lengths = [5, 7, 3, 9, 4]
sh = shard_tracks(lengths, world)
assert sh[0][0] == 0 and sh[-1][1] == len(lengths) and sh[-1][3] == sum(lengths)
t0, t1, f0, f1 = sh[rank]
full = torch.arange(sum(lengths) * 32, dtype=torch.int64).reshape(-1, 32).to(torch.uint8)
ex = ScoreExchange(sum(lengths), None, 0, None, 'nccl')
got, _ = ex.gather(full[f0:f1].clone(), f0, None, sizes=[s[3] - s[2] for s in sh])
assert torch.equal(got, full)
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


def test_all_gather_rows_gloo_world2(tmp_path):
    """N>1 path on CPU: two ranks, ragged and equal shards, gathered rows in frame order."""
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER.replace('{root!r}', repr(ROOT)))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29571', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '3'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'frames/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['value'] == line['value']


def test_report_writers_known_answers(tmp_path):
    """report.py (base.py:329-397 text files) without the reference tree: exact cell text and row layout."""
    import csv
    from poserisk_release_b200 import report
    e = np.array([[[1.23456, -0.0004, 179.9996], [0.0, 10.0, -20.5]]])
    assert report.pose_to_str(e) == [['(1.235, -0.000, 180.000)', '(0.000, 10.000, -20.500)']]
    ts = (0, np.array([1, 3]), 5)
    results = [{'score': np.int64(4), 'log_score': [2, 1, 1, '2,3']}, {'score': np.int64(7), 'log_score': [3, 2, 1, '4,4']}]
    final, scores, logs = report.post_processing(results)
    assert final[0] == 5.5 and final[3] == 7 and final[4] == 4 and np.isnan(final[2]) and list(scores) == [4, 7]
    report.save_csv([['a', 'b']] * 2, ts, scores, ['Trunk', 'Neck', 'Leg', 'Arm'], logs,
                    [{'k1': 'x', 'k2': 'y'}, {'k1': 'z', 'k2': 'w'}], str(tmp_path), title='REBA')
    rows = list(csv.reader(open(tmp_path / 'REBA_score_log.csv')))
    assert rows[0] == ['Frame', 'Final_score', 'Joint Score', 'Trunk', 'Neck', 'Leg', 'Arm']
    assert rows[1] == ['0'] and rows[2] == ['1', '4', '', '2', '1', '1', '2,3'] and rows[4] == ['3', '7', '', '3', '2', '1', '4,4']
    assert len(rows) == 6
    rows = list(csv.reader(open(tmp_path / 'REBA_eval_pose_log.csv')))
    assert rows[0] == ['Frame', '', 'k1', 'k2'] and rows[2] == ['1', '', 'x', 'y']
    report.save_csv_pose_log([['p0', 'p1', 'p2']] * 2, ts, str(tmp_path), ['Neck'], ['PELVIS', 'NECK', 'HEAD'])
    rows = list(csv.reader(open(tmp_path / 'pose_log.csv')))
    assert rows[0] == ['Frame', 'Joint Pose', 'Neck'] and rows[2] == ['1', '', 'p1']
    txt = report.result_text((5.5, 7.0, float('nan'), 7, 4), 3, 'Medium risk.', 'RULA')
    assert txt.startswith('AVG Score: 5.5 \n%50 Score: 7.0 \n%10 Score: nan ') and txt.endswith('Action: Medium risk.')


_COMPAT_PROBE = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import poserisk_release_b200
sys.path.insert(0, os.path.join(os.path.dirname(poserisk_release_b200.__file__), 'compat'))   # INTEGRATION.md section 1
# exactly the bare imports of lib/core/base.py:24-31 and lib/utils/smpl.py:5
from smpl import SMPL
from coord_utils import axis_angle_to_euler_angle, rot_to_angle, get_joint_cam
from reba import REBA
from rula import RULA
from smplpytorch.pytorch.smpl_layer import SMPL_Layer
mods = {{c.__name__: c.__module__ for c in (SMPL, REBA, RULA, SMPL_Layer, axis_angle_to_euler_angle, rot_to_angle, get_joint_cam)}}
m = SMPL()
assert isinstance(m.layer['female'], SMPL_Layer)
print(json.dumps({{'mods': mods, 'model_path': m.model_path, 'layers': sorted(m.layer), 'jr': list(m.joint_regressor.shape)}}))
'''


def test_compat_shims_resolve_the_reference_imports(tmp_path):
    """The drop-in mechanism itself: with compat/ first on sys.path the bare module names that base.py:24-31 and
    smpl.py:5 import resolve to this package, and SMPL() builds its three layers through them."""
    script = tmp_path / 'probe.py'
    script.write_text(_COMPAT_PROBE.format(root=ROOT))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, PRK_SYNTHETIC_SMPL='1'), cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    assert all(v.startswith('poserisk_release_b200.') for v in got['mods'].values()), got['mods']
    assert got['layers'] == ['female', 'male', 'neutral'] and got['jr'] == [29, 6890]
    assert got['model_path'] == os.path.join('data', 'base_data', 'human_models')


def test_missing_model_file_raises_unless_synthetic_is_requested(monkeypatch):
    """serialization.py:10 open()s the .pkl: a wrong model_root must not silently give a random model."""
    from poserisk_release_b200 import SMPL_Layer, model_provider
    monkeypatch.delenv('PRK_SYNTHETIC_SMPL', raising=False)
    with pytest.raises(FileNotFoundError, match='SMPL_NEUTRAL.pkl'):
        SMPL_Layer(model_root='no/such/dir')
    with pytest.raises(FileNotFoundError):
        model_provider.get_model_data('male', None)
    with pytest.warns(UserWarning, match='SYNTHETIC'):
        lay = SMPL_Layer(model_root='no/such/dir', allow_synthetic=True)
    assert lay.smpl_data.synthetic
    lay = SMPL_Layer(model_data=model_provider.synthetic_smpl('female'), gender='female')      # explicit data: no warning needed
    assert lay.smpl_data.gender == 'female'


def test_action_levels_known_answers():
    from poserisk_release_b200 import REBA, RULA
    r, u = REBA(), RULA()
    assert [r.action_level(s)[0] for s in (1, 2, 3, 4, 7, 8, 10, 11, 15)] == [1, 2, 2, 3, 3, 4, 4, 5, 5]
    assert [u.action_level(s)[0] for s in (1, 2, 3, 4, 5, 6, 7, 9)] == [1, 1, 2, 2, 3, 3, 4, 4]
    assert r.action_level(0) == (None, None) and u.action_level(-3) == (None, None)
    assert r.action_level(7.4) == (3, "Medium risk. Further Investigate. Change Soon.")
    assert u.action_level(6.5)[0] == 3 and u.action_level(6.51)[0] == 4        # round-half-even, like the reference


def test_save_obj_known_answer(tmp_path):
    from poserisk_release_b200 import report
    v = np.array([[0.5, -1.25, 3.0], [1e-7, 2.0, 1234.5678]], np.float32)
    f = np.array([[0, 1, 1]], np.uint32)
    report.save_obj(v, f, str(tmp_path / 'm.obj'))
    assert (tmp_path / 'm.obj').read_text() == 'v 0.5 -1.25 3.0\nv 1e-07 2.0 1234.5677\nf 1/1 2/2 2/2\n'
    report.save_obj(v[:1].astype(np.float64), None, str(tmp_path / 'n.obj'))
    assert (tmp_path / 'n.obj').read_text() == 'v 0.5 -1.25 3.0\n'


def test_vertex_kernel_unit_ranges_cover_every_unit_once():
    """The vertex kernel's CTAs (pairs) take contiguous unit ranges balanced by COST (a frame-tile switch inside a range is
    charged switch_cost16 / 16 units, csrc/prk_internal.h fused_first_unit).  Host arithmetic through the verification hook:
    for every batch shape the ranges are contiguous, cover [0, n_units) exactly once, and no range's cost exceeds the mean by
    more than one unit plus one switch."""
    import ctypes as C
    from poserisk_release_b200 import _lib
    L = _lib.lib()
    NT = 216
    for tiles in (1, 2, 3, 16, 17, 64, 256):
        for n_ranges in (1, 2, 37, 74, 148):
            for cost16 in (0, 16, 48, 200):
                n_units = tiles * NT
                prev_end, worst = 0, 0.0
                total = 16 * n_units + cost16 * (tiles - 1)
                for k in range(n_ranges):
                    u0, u1 = C.c_int64(), C.c_int64()
                    _lib.check(L.prk_debug_unit_range(n_units, n_ranges, cost16, k, C.byref(u0), C.byref(u1)))
                    assert u0.value == prev_end and u1.value >= u0.value, (tiles, n_ranges, cost16, k)
                    prev_end = u1.value
                    # switches strictly inside the range: tile boundaries b with u0 < b < u1 ... plus one at u0 if u0 is a boundary > 0
                    inside = sum(1 for b in range(NT, n_units, NT) if u0.value < b < u1.value)
                    worst = max(worst, 16 * (u1.value - u0.value) + cost16 * inside)
                assert prev_end == n_units
                assert worst <= total / n_ranges + 16 + cost16 + 16, (tiles, n_ranges, cost16, worst, total / n_ranges)
    bad = C.c_int64()
    assert L.prk_debug_unit_range(100, 4, 0, 0, C.byref(bad), C.byref(bad)) != 0       # not a multiple of 216


def _e4m3(x):
    """Round-to-nearest-even quantisation to OCP e4m3 (4 significant bits, normal exponents -6..8, subnormal step 2^-9,
    saturating at 448) -- the format of the blend kernel's cross-term operands."""
    x = np.asarray(x, np.float64)
    a = np.minimum(np.abs(x), 448.0)
    e = np.floor(np.log2(np.maximum(a, 2.0 ** -20)))
    e = np.clip(e, -6, 8)
    q = 2.0 ** (e - 3)                      # spacing: 3 explicit mantissa bits
    return np.sign(x) * np.round(a / q) * q  # numpy rounds half to even


def test_blend_operand_scheme_numpy_model():
    """The arithmetic of the vertex kernel's blend stage restated in numpy (csrc/prk_internal.h "K12 operand layout"):
    fp16 main product + two e4m3 cross terms for the pose blend shapes, two fp16 parts per factor for betas x shapedirs,
    the template as 2^15 (u1 + u2 + u3), everything scaled by 2^S, fp32 accumulation.  Against float64 the model must stay
    far inside north_star's 1e-5 (it is what the GPU parity tests then measure on the device: 3e-7 .. 5e-7)."""
    from poserisk_release_b200.model_provider import synthetic_smpl
    from scipy.spatial.transform import Rotation
    m = synthetic_smpl('neutral')
    pd = m.posedirs.reshape(-1, 207).astype(np.float64)[::7]          # a seventh of the vertex coordinates keeps it quick
    sd = m.shapedirs.reshape(-1, 10).astype(np.float64)[::7]
    vt = m.v_template.reshape(-1).astype(np.float64)[::7]
    S = int(np.floor(np.log2(min(16384 / np.abs(m.posedirs).max(), 32768 / np.abs(m.shapedirs).max(),
                                 2.0 ** 30 / np.abs(m.v_template).max()))))
    f16 = lambda a: np.asarray(a, np.float64).astype(np.float16).astype(np.float64)
    rng = np.random.default_rng(0)
    for pose_s, beta_s in ((0.35, 1.0), (0.6, 3.0), (1.2, 10.0)):
        B = 64
        rv = rng.standard_normal((B, 23, 3)) * pose_s
        R = Rotation.from_rotvec(rv.reshape(-1, 3)).as_matrix().reshape(B, 23, 9).astype(np.float32).astype(np.float64)
        F = (R - np.eye(3).reshape(1, 1, 9)).reshape(B, 207)
        beta = (rng.standard_normal((B, 10)) * beta_s).astype(np.float32).astype(np.float64)
        ref = vt[None] + beta @ sd.T + F @ pd.T
        P, Sd, T = np.ldexp(pd, S), np.ldexp(sd, S), np.ldexp(vt, S)
        Fh, Ph = f16(F), f16(P)
        acc = Fh @ Ph.T                                                 # 13 kind::f16 MMAs
        acc += _e4m3((F - Fh) * 4096.0) @ _e4m3(P / 4096.0).T + _e4m3(F) @ _e4m3(P - Ph).T     # 13 kind::f8f6f4 MMAs
        bh, sh = f16(beta), f16(Sd)
        bl, sl = f16(beta - bh), f16(Sd - sh)
        u = T * 2.0 ** -15
        u1 = f16(u); u2 = f16(u - u1); u3 = f16(u - u1 - u2)
        acc += bh @ sh.T + bl @ sh.T + bh @ sl.T + 2.0 ** 15 * (u1 + u2 + u3)[None]             # 3 kind::f16 MMAs
        out = np.ldexp(acc.astype(np.float32).astype(np.float64), -S)
        err = np.abs(out - ref).max() / np.abs(ref).max()
        assert err < 1e-6, (pose_s, beta_s, err)
    # the format emulation itself: exact e4m3 values survive, the spacing above 16 is 2, saturation at 448
    assert np.array_equal(_e4m3([0.0, 1.0, 1.125, 17.0, 18.9, 500.0, -0.001953125]), [0.0, 1.0, 1.125, 16.0, 18.0, 448.0, -0.001953125])
