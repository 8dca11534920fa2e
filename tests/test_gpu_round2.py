"""GPU tests added in round 2: ABI v2 behaviour (track-id bounds, debug sequences through the pipeline call,
back-to-back host calls of different sizes), the peer-memory all-gather, config 4 at BASELINE.json's size,
the NaN-joint semantics and a race stress test of the vertex kernel.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import oracle
from poserisk_release_b200.model_provider import synthetic_smpl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5
from test_gpu_parity import EXAMPLE_INFO, relerr, same_records  # noqa: E402


@pytest.fixture(scope='module')
def engine():
    from poserisk_release_b200 import PoseRiskEngine
    return PoseRiskEngine('cuda:0', genders=('neutral', 'female', 'male'))


def make(B, seed, scale=0.4):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 72, generator=g) * scale, torch.randn(B, 10, generator=g), torch.randn(B, 3, generator=g) * 0.1


# --------------------------------------------------------------------------- ABI v2
def test_out_of_range_track_ids_are_flagged_not_dereferenced(engine):
    """A track id outside [0, n_tracks) is an IndexError in the reference (a Python list index); here the frame is
    scored with track 0's additional information and carries flags bit 1 -- in both scoring kernels."""
    from golden.make_golden import random_info
    from poserisk_release_b200 import _runtime
    rng = np.random.default_rng(3)
    infos = [random_info(rng) for _ in range(3)]
    for B in (257, 140001):                                  # lanes kernel / thread-per-frame kernel
        pose, _, _ = make(B, 5)
        track = rng.integers(0, 3, B).astype(np.int32)
        bad = rng.choice(B, 40, replace=False)
        track_bad = track.copy()
        track_bad[bad[:20]] = -1 - rng.integers(0, 1000, 20)
        track_bad[bad[20:]] = 3 + rng.integers(0, 1 << 20, 20)
        out = engine.run(pose.cuda(), add_info=infos, track_of_frame=track_bad, want_verts=False)
        rec = _runtime.records_to_numpy(out['scores'])
        expect_track = track.copy()
        expect_track[bad] = 0
        ref = oracle.score_pose(pose.numpy(), infos, expect_track)
        assert same_records(rec, ref).all()
        flagged = (rec['flags'] & 2) != 0
        assert flagged.sum() == 40 and flagged[bad].all()
    with pytest.raises(IndexError):
        engine.run_tracks(pose[:10].cuda(), None, None, infos, np.full(10, 5, np.int32), ['male', 'female', 'neutral'])


def test_pipeline_call_emits_debug_euler_sequences(engine):
    """engine.run(debug_joints=...) = one prk_pipeline call: joints + scores + the --debug_joints Euler
    sequences (base.py:144-146), for the staged (<= 6 joints) and the direct-store (> 6 joints) path, ragged batch
    sizes, float32 poses.  Angles against the oracle; scores and joints identical to a run without debug output."""
    from poserisk_release_b200 import _runtime
    for B, ids in ((1, [12]), (33, [12, 16, 17, 3]), (4097, [0, 23, 5]), (2500, list(range(0, 24, 3))), (140003, [12, 16, 17, 3])):
        pose, betas, trans = make(B, 100 + B % 97, 0.6)
        out = engine.run(pose.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO, want_verts=False, debug_joints=ids)
        plain = engine.run(pose.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO, want_verts=False)
        assert torch.equal(out['scores'], plain['scores']) and torch.equal(out['joints'], plain['joints'])
        n_ref = min(B, 3000)
        rec_ref, eul_ref = oracle.score_pose(pose[:n_ref].numpy(), EXAMPLE_INFO, want_euler=True)
        eul_ref = np.asarray(eul_ref).reshape(n_ref, 24, 3)[:, ids]
        got = out['euler'].cpu().numpy()
        assert got.shape == (B, len(ids), 3)
        assert np.abs(got[:n_ref] - eul_ref).max() < 1e-9
        assert same_records(_runtime.records_to_numpy(out['scores'][:n_ref]), rec_ref).all()
        # the stand-alone scoring entry point gives the same bits
        _, eul2 = engine.euler_debug(pose.cuda(), ids, EXAMPLE_INFO)
        assert torch.equal(eul2, out['euler'])


def test_float64_pose_scoring_big_batch(engine):
    """The float64 instantiation of the staged thread-per-frame scoring kernel (576-byte rows)."""
    from poserisk_release_b200 import _runtime
    B = 3001
    pose = (torch.randn(B, 72, generator=torch.Generator().manual_seed(8), dtype=torch.float64) * 0.7)
    scores, eul = engine.euler_debug(pose.cuda(), [3, 12, 16, 17, 20], EXAMPLE_INFO)
    rec_ref, eul_ref = oracle.score_pose(pose.numpy(), EXAMPLE_INFO, want_euler=True)
    eul_ref = np.asarray(eul_ref).reshape(B, 24, 3)[:, [3, 12, 16, 17, 20]]
    assert np.abs(eul.cpu().numpy() - eul_ref).max() < 1e-9
    assert same_records(_runtime.records_to_numpy(scores), rec_ref).all()


def test_back_to_back_host_calls_of_different_sizes(engine):
    """the staging layout of prk_pipeline_host depends on B, so a call may only run ahead of the
    previous one when both use the same layout.  B1 -> B2 -> B1 on the same workspace without any synchronisation in
    between, with two model handles interleaved, must give what synchronised calls give."""
    sizes = [4096, 1000, 4096, 37, 2048, 2048, 4096]
    genders = ['neutral', 'neutral', 'female', 'neutral', 'male', 'male', 'neutral']
    ins, outs = [], []
    for k, (B, gd) in enumerate(zip(sizes, genders)):
        pose, betas, trans = make(B, 900 + k)
        ins.append(tuple(t.pin_memory() for t in (pose, betas, trans)))
        outs.append((torch.empty(B, 24, 3).pin_memory(), torch.empty(B, 32, dtype=torch.uint8).pin_memory(),
                     torch.empty(B, 6890, 3, device='cuda')))
    torch.cuda.synchronize()
    for _ in range(3):                                       # a few rounds: set parity of the staging sets varies
        for (p, b, t), (j, s, v), gd in zip(ins, outs, genders):
            engine.run_host(p, b, t, EXAMPLE_INFO, None, j, s, gender=gd, verts_out=v)
    torch.cuda.synchronize()
    for (p, b, t), (j, s, v), gd in zip(ins, outs, genders):
        ref = engine.run(p.cuda(), b.cuda(), t.cuda(), add_info=EXAMPLE_INFO, gender=gd)
        torch.cuda.synchronize()
        assert torch.equal(ref['verts'], v) and torch.equal(ref['joints'].cpu(), j) and torch.equal(ref['scores'].cpu(), s)


def test_refused_pipeline_call_launches_nothing(engine):
    """Arguments and workspace are checked before the first launch: a NULL pose / too-small workspace
    returns an error, leaves the launch counter and the output untouched, and the context stays usable."""
    from poserisk_release_b200 import _lib, _runtime
    L = _lib.lib()
    h = engine.models['neutral']
    B = 64
    pose, _, _ = make(B, 1)
    pose = pose.cuda()
    info = _runtime.addinfo_tensor(EXAMPLE_INFO, torch.device('cuda:0'))
    scores = torch.full((B, 32), 7, dtype=torch.uint8, device='cuda')
    joints = torch.zeros(B, 24, 3, device='cuda')
    verts = torch.zeros(B, 6890, 3, device='cuda')
    ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    wptr = C.c_void_p(ws.data_ptr() + (-ws.data_ptr()) % 1024)
    stream = _runtime.stream_ptr(torch.device('cuda:0'))
    before = _lib.launch_count()

    def call(pose_ptr, ws_bytes, n_tracks=1):
        return L.prk_pipeline(h.handle, pose_ptr, None, None, -1, _runtime.ptr(info), n_tracks, None, B, _runtime.ptr(verts), 0,
                              _runtime.ptr(joints), _runtime.ptr(scores), None, None, 0, None, None, 0, wptr, ws_bytes, stream)
    assert call(None, 1 << 19) == 1                          # PRK_ERR_INVALID_ARG
    assert call(_runtime.ptr(pose), 4096) == 3               # PRK_ERR_WORKSPACE
    assert call(_runtime.ptr(pose), 1 << 19, n_tracks=0) == 1
    torch.cuda.synchronize()
    assert _lib.launch_count() == before and bool((scores == 7).all())
    assert call(_runtime.ptr(pose), 1 << 19) == 0
    torch.cuda.synchronize()
    assert _lib.launch_count() == before + 3


# --------------------------------------------------------------------------- multi-person at scale
def test_config4_sixteen_tracks_of_ten_thousand_frames(engine):
    """BASELINE.json config 4 at its stated size: 16 tracks x 10,000 frames, genders cycling male/female/neutral,
    per-track additional information.  Scores of every frame against the oracle, joints of every frame, vertices
    on a sample of each track; no gather / scatter: the run count equals the number of gender changes."""
    from golden.make_golden import random_info
    from poserisk_release_b200 import _runtime
    rng = np.random.default_rng(44)
    T, F = 16, 10000
    genders = [('male', 'female', 'neutral')[t % 3] for t in range(T)]
    infos = [random_info(rng) for _ in range(T)]
    track = np.repeat(np.arange(T), F).astype(np.int32)
    pose, betas, trans = make(T * F, 45, 0.45)
    joints = torch.empty(T * F, 24, 3, device='cuda')
    scores = torch.empty(T * F, 32, dtype=torch.uint8, device='cuda')
    # vertices of 160k frames are 13 GB: fine on a 180 GB part, and the point of the test
    out = engine.run_tracks(pose.cuda(), betas.cuda(), trans.cuda(), infos, track, genders, joints_out=joints, scores_out=scores)
    torch.cuda.synchronize()
    assert out['runs'] == T
    ref = oracle.score_pose(pose.numpy(), infos, track)
    assert same_records(_runtime.records_to_numpy(out['scores']), ref).all()
    for t in range(T):
        sl = slice(t * F + 4990, t * F + 5010)               # straddles nothing, but sits deep inside a 65,536-frame chunk
        v_ref, j_ref = oracle.smpl_forward(synthetic_smpl(genders[t]), pose[sl].numpy(), betas[sl].numpy(), trans[sl].numpy())
        assert relerr(out['verts'][sl].cpu().numpy(), v_ref) < TOL
        assert relerr(out['joints'][sl].cpu().numpy(), j_ref) < TOL
    _, j_all = oracle.smpl_forward(synthetic_smpl('male'), pose[:F].numpy(), betas[:F].numpy(), trans[:F].numpy(), want_verts=False)
    assert relerr(out['joints'][:F].cpu().numpy(), j_all) < TOL
    del out
    torch.cuda.empty_cache()


# --------------------------------------------------------------------------- NaN semantics
def test_nan_pose_poisons_its_frame_and_nothing_else(engine):
    """A non-finite pose entry: in the reference the joint's R - I features are NaN, so v_posed -- a dense product
    over all 207 features (smpl_layer.py:97-99) -- and with it EVERY vertex of that frame is NaN, the chain below
    the joint is NaN (:109-119), and axis_angle_to_euler_angle stops at assert(isRotationMatrix(R))
    (coord_utils.py:70).  Same here: the whole frame's mesh is NaN (the NaN enters through the tcgen05 blend GEMM's
    A' row, i.e. one accumulator lane), joints at and below the joint are NaN, the record carries flags bit 0 -- and
    every other frame of the batch, including the 127 that share its MMA tile, keeps its bits."""
    from poserisk_release_b200 import _runtime
    B = 300
    pose, betas, trans = make(B, 71)
    clean = engine.run(pose.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO)
    f_bad, j_bad = 137, 18                                   # L_Elbow: descendants 20 (L_Wrist), 22 (L_Hand)
    pose_bad = pose.clone()
    pose_bad[f_bad, j_bad * 3 + 1] = float('nan')
    out = engine.run(pose_bad.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO)
    v, vc = out['verts'].cpu().numpy(), clean['verts'].cpu().numpy()
    keep = np.arange(B) != f_bad
    assert np.array_equal(v[keep], vc[keep]) and torch.equal(out['joints'][keep], clean['joints'][keep])
    assert np.isnan(v[f_bad]).all()
    j = out['joints'][f_bad].cpu().numpy()
    assert np.isnan(j[[20, 22]]).all() and not np.isnan(j[:18]).any()   # the chain below the joint
    rec = _runtime.records_to_numpy(out['scores'])
    assert rec['flags'][f_bad] & 1 and not (rec['flags'][keep] & 1).any()
    assert same_records(rec[keep], _runtime.records_to_numpy(clean['scores'])[keep]).all()
    # root joint: not a pose feature, but every transform of the chain descends from it
    pose_bad = pose.clone()
    pose_bad[5, 0] = float('inf')
    out = engine.run(pose_bad.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO)
    assert torch.isnan(out['verts'][5]).all() and torch.isnan(out['joints'][5][1:]).all()
    assert torch.equal(out['joints'][5][0], clean['joints'][5][0])     # the root's position is its rest joint + trans
    assert torch.equal(out['verts'][6:], clean['verts'][6:]) and torch.equal(out['verts'][:5], clean['verts'][:5])


# --------------------------------------------------------------------------- peer-memory exchange
def test_peer_allgather_two_ranks_in_one_process():
    """prk_allgather_rows with world = 2 on one device: two communicators in this process (same-process peers are
    addressed directly, no CUDA IPC), one stream each.  Ragged shards, 32-byte records and 24-byte-per-joint Euler
    rows (8-byte units), many epochs (both slots, flag reuse), and the gathered view of each rank."""
    from poserisk_release_b200 import _runtime
    dev = torch.device('cuda:0')
    n = 10007
    from poserisk_release_b200 import _lib
    L = _lib.lib()
    comms = []
    for r in range(2):
        h = C.c_void_p()
        _lib.check(L.prk_comm_create(C.byref(h), r, 2, 0, n * 72))
        comms.append(h)
    hb = int(L.prk_comm_handle_bytes())
    blob = b''
    for h in comms:
        buf = C.create_string_buffer(hb)
        _lib.check(L.prk_comm_get_handle(h, buf))
        blob += buf.raw
    for h in comms:
        _lib.check(L.prk_comm_open_peers(h, blob))
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    cut = 4001                                               # rank 0: rows [0, cut), rank 1: [cut, n)
    try:
        for epoch in range(7):
            for row_shape, dtype in (((32,), torch.uint8), ((3, 3), torch.float64)):
                full = torch.randint(0, 255, (n,) + row_shape, device=dev).to(dtype)
                outs = []
                row_bytes = full[0].numel() * full.element_size()
                torch.cuda.synchronize()
                for r, (lo, hi) in enumerate(((0, cut), (cut, n))):
                    local = full[lo:hi].contiguous()
                    g = C.c_void_p()
                    with torch.cuda.stream(streams[r]):
                        _lib.check(L.prk_allgather_rows(comms[r], _runtime.ptr(local), hi - lo, lo, row_bytes, C.byref(g),
                                                        C.c_void_p(streams[r].cuda_stream)))
                    outs.append((g.value, local))
                torch.cuda.synchronize()
                for r in range(2):
                    assert L.prk_comm_status(comms[r]) == 0
                    view = _runtime._DevView(outs[r][0], full.numel(), dtype, dev, None).tensor.view(full.shape)
                    assert torch.equal(view, full), (epoch, r)
                    assert int(L.prk_comm_gathered(comms[r])) == outs[r][0]
    finally:
        torch.cuda.synchronize()
        for h in comms:
            L.prk_comm_destroy(h)


def test_exchange_object_single_rank(engine):
    """world = 1: the exchange inside prk_pipeline degenerates to a copy into the rank's own buffer; the gathered views
    equal the call's outputs (scores and debug Euler rows), for the device and the host entry point."""
    from poserisk_release_b200.distributed import ScoreExchange
    B = 5000
    pose, betas, trans = make(B, 12)
    ex = ScoreExchange(B, torch.device('cuda:0'), n_debug=4)
    assert ex.used == 'peer'
    out = engine.run(pose.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO, want_verts=False,
                     debug_joints=[12, 16, 17, 3], exchange=ex, frame_offset=0)
    s, e = ex.collect(out['scores'], 0, out['euler'])
    torch.cuda.synchronize()
    assert torch.equal(s, out['scores']) and torch.equal(e, out['euler'])
    ex.check()
    j = torch.empty(B, 24, 3).pin_memory()
    sc = torch.empty(B, 32, dtype=torch.uint8).pin_memory()
    ex2 = ScoreExchange(B, torch.device('cuda:0'))
    engine.run_host(pose.pin_memory(), betas.pin_memory(), trans.pin_memory(), EXAMPLE_INFO, None, j, sc, exchange=ex2)
    s2, _ = ex2.collect(None, 0)
    torch.cuda.synchronize()
    assert torch.equal(s2.cpu(), sc) and torch.equal(sc, out['scores'].cpu())


_WORLD2 = r'''
import os, sys
sys.path.insert(0, {root!r})
os.environ.setdefault('PRK_SYNTHETIC_SMPL', '1')
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
from poserisk_release_b200 import PoseRiskEngine
from poserisk_release_b200.distributed import run_sharded, ScoreExchange
eng = PoseRiskEngine(dev)
info = {info!r}
n = 100003
g = torch.Generator().manual_seed(5)
pose = torch.randn(n, 72, generator=g) * 0.4
ids = [12, 16, 17, 3]
whole = eng.run(pose.to(dev), add_info=info, want_verts=False, debug_joints=ids)
for transport in ('peer', 'nccl'):
    out = run_sharded(eng, pose.to(dev), None, None, info, debug_joints=ids, transport=transport)
    torch.cuda.synchronize()
    assert out['transport'] == transport, out['transport']
    assert torch.equal(out['scores'], whole['scores']) and torch.equal(out['euler'], whole['euler']), transport
# many exchanges back to back (slot reuse across processes)
ex = ScoreExchange(n, dev, 0, None, 'peer')
lo, hi = out['range']
for k in range(20):
    local = (whole['scores'][lo:hi] + k).contiguous()
    got, _ = ex.gather(local, lo)
    torch.cuda.synchronize()
    assert torch.equal(got, whole['scores'] + k), k
ex.check()
dist.barrier()
dist.destroy_process_group()
print('ok', rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_sharded_run_equals_single_gpu_run_world2(tmp_path):
    """Two processes, two GPUs: scores and debug Euler sequences of a sharded run (ragged shards) are bit-identical to
    one GPU's, over the peer-memory exchange (CUDA IPC between the processes) and over NCCL."""
    script = tmp_path / 'w2.py'
    script.write_text(_WORLD2.replace('{root!r}', repr(ROOT)).replace('{info!r}', repr(EXAMPLE_INFO)))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29577', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


# --------------------------------------------------------------------------- race stress of the vertex kernel
def test_vertex_kernel_stress_ragged_batches_twice_into_poisoned_buffers(engine):
    """Race evidence for the fused kernel's ~70 mbarriers, deferred stores and split-phase hand-overs (no
    compute-sanitizer on the GPU pool): many ragged batch sizes, each run TWICE into NaN-filled outputs on different
    streams with other work in flight, compared bit for bit with each other; one vertex row of each is checked against
    the oracle; and the dense-weights (> 4 non-zeros, kGroups = 0) instantiation at 4096 + 77 frames."""
    from poserisk_release_b200 import PoseRiskEngine, SMPL_Layer
    from poserisk_release_b200.model_provider import SMPLModelData
    sizes = [1, 31, 127, 128, 129, 255, 300, 511, 1000, 1025, 2047, 4095, 4096, 4097, 4173, 6000, 8191, 9473]
    side = torch.cuda.Stream()
    noise = torch.empty(64 << 20, device='cuda')
    m = synthetic_smpl('neutral')
    for k, B in enumerate(sizes):
        pose, betas, trans = make(B, 300 + k, 0.5)
        p, b, t = pose.cuda(), betas.cuda(), trans.cuda()
        res = []
        for rep in range(2):
            v = torch.full((B, 6890, 3), float('nan'), device='cuda')
            with torch.cuda.stream(side):                    # unrelated traffic on another stream
                noise.normal_()
            out = engine.run(p, b, t, add_info=EXAMPLE_INFO, verts_out=v)
            res.append((v, out['joints'].clone(), out['scores'].clone()))
        torch.cuda.synchronize()
        assert not torch.isnan(res[0][0]).any(), B
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2]), B
        pick = [0, B // 2, B - 1]
        v_ref, _ = oracle.smpl_forward(m, pose[pick].numpy(), betas[pick].numpy(), trans[pick].numpy())
        assert relerr(res[0][0][pick].cpu().numpy(), v_ref) < TOL, B
        del res
    # dense weights: every vertex bound to 6 joints -> the run-time group loop (kGroups = 0)
    rng = np.random.default_rng(17)
    w = np.zeros((6890, 24), np.float32)
    for vtx in range(6890):
        idx = rng.choice(24, 6, replace=False)
        x = rng.uniform(0.05, 1.0, 6)
        w[vtx, idx] = x / x.sum()
    md = SMPLModelData(**{**m.__dict__, 'weights': w})
    lay = SMPL_Layer(model_data=md)
    B = 4096 + 77
    pose, betas, trans = make(B, 999, 0.5)
    runs = [lay(pose.cuda(), betas.cuda(), trans.cuda()) for _ in range(2)]
    torch.cuda.synchronize()
    assert torch.equal(runs[0][0], runs[1][0]) and not torch.isnan(runs[0][0]).any()
    pick = [0, 2048, 4096, B - 1]
    v_ref, j_ref = oracle.smpl_forward(md, pose[pick].numpy(), betas[pick].numpy(), trans[pick].numpy())
    assert relerr(runs[0][0][pick].cpu().numpy(), v_ref) < TOL and relerr(runs[0][1][pick].cpu().numpy(), j_ref) < TOL


# --------------------------------------------------------------------------- aligned vertex rows (bulk tensor stores)
def test_aligned_vertex_rows_equal_the_dense_layout_bit_for_bit(engine):
    """The engine's own vertex tensors have rows padded to 16 bytes (pitch 20672 floats) and are written by bulk tensor
    stores from warp-private tiles; a dense (B, 6890, 3) tensor goes through the shared staging tile and 8-byte stores.
    Both must hold identical bits -- ragged batches, CTA pairs and single CTAs, slices of a larger aligned buffer (config 4's
    per-track slices).  The TMA unit clips a box at the next 16-byte boundary behind the tensor's last column, so the two pad
    floats that complete a row to 16 bytes may be overwritten (they belong to nobody); anything behind them must survive."""
    from poserisk_release_b200 import _lib, _runtime
    for B in (1, 2, 31, 128, 129, 300, 641, 1500, 4096, 4173):
        pose, betas, trans = make(B, 700 + B % 89, 0.5)
        p, b, t = pose.cuda(), betas.cuda(), trans.cuda()
        pitch = _lib.VERTS_PITCH_ALIGNED + 8 * (B % 2)             # 20672, and a wider pitch with a guard band behind the pad
        flat = torch.full((B, pitch), float('nan'), device='cuda')
        flat[:, 20672:] = 12345.0                                 # behind the 16-byte pad: must survive
        v_al = flat[:, :20670].view(B, 6890, 3)
        assert _runtime.verts_pitch(v_al) in (20670, pitch) and v_al.data_ptr() % 16 == 0
        dense = torch.full((B, 6890, 3), float('nan'), device='cuda')
        engine.run(p, b, t, add_info=EXAMPLE_INFO, verts_out=v_al)
        engine.run(p, b, t, add_info=EXAMPLE_INFO, verts_out=dense)
        auto = engine.run(p, b, t, add_info=EXAMPLE_INFO)['verts']     # engine-allocated: aligned
        torch.cuda.synchronize()
        assert not torch.isnan(dense).any() and torch.equal(v_al, dense) and torch.equal(auto, dense), B
        assert bool((flat[:, 20672:] == 12345.0).all()), B
        if B > 1:
            assert auto.stride(0) == _lib.VERTS_PITCH_ALIGNED and not auto.is_contiguous()
        assert torch.equal(auto.contiguous(), dense)
    # slices of an aligned buffer at odd and even frame offsets (what run_tracks hands to each gender's model)
    B = 900
    pose, betas, trans = make(B, 31, 0.5)
    big = _runtime.aligned_verts(B, torch.device('cuda:0'))
    ref = engine.run(pose.cuda(), betas.cuda(), trans.cuda(), add_info=EXAMPLE_INFO, verts_out=torch.empty(B, 6890, 3, device='cuda'))['verts']
    for lo, hi in ((0, 301), (301, 555), (555, 900)):
        engine.run(pose[lo:hi].cuda(), betas[lo:hi].cuda(), trans[lo:hi].cuda(), add_info=EXAMPLE_INFO, verts_out=big[lo:hi])
    torch.cuda.synchronize()
    assert torch.equal(big, ref)
    # the host-buffer entry point with aligned rows
    j = torch.empty(B, 24, 3).pin_memory()
    sc = torch.empty(B, 32, dtype=torch.uint8).pin_memory()
    v_host_path = _runtime.aligned_verts(B, torch.device('cuda:0'))
    engine.run_host(pose.pin_memory(), betas.pin_memory(), trans.pin_memory(), EXAMPLE_INFO, None, j, sc, verts_out=v_host_path)
    torch.cuda.synchronize()
    assert torch.equal(v_host_path, ref)


def test_aligned_rows_through_the_drop_in_layer_and_the_oracle():
    """SMPL_Layer(aligned_verts=True): same values as the dense default, against the oracle; the C ABI refuses a pitch that
    cannot be honoured."""
    from poserisk_release_b200 import SMPL_Layer, _lib, _runtime
    from poserisk_release_b200.model_provider import SMPLModelData
    B = 777
    pose, betas, trans = make(B, 5, 0.5)
    lay_a, lay_d = SMPL_Layer(aligned_verts=True), SMPL_Layer()
    va, ja = lay_a(pose.cuda(), betas.cuda(), trans.cuda())
    vd, jd = lay_d(pose.cuda(), betas.cuda(), trans.cuda())
    assert vd.is_contiguous() and not va.is_contiguous() and tuple(va.shape) == (B, 6890, 3)
    assert torch.equal(va, vd) and torch.equal(ja, jd)
    v_ref, j_ref = oracle.smpl_forward(synthetic_smpl('neutral'), pose.numpy(), betas.numpy(), trans.numpy())
    assert relerr(va.cpu().numpy(), v_ref) < TOL and relerr(ja.cpu().numpy(), j_ref) < TOL
    # a pitch that is not a multiple of 4 floats, or smaller than a row, is an invalid argument; a padded pitch with a
    # dense-weights model (run-time group loop, plain stores only) is unsupported
    L = _lib.lib()
    h = lay_d._handle(torch.device('cuda:0'))
    ws, ws_bytes, _keep = _runtime.workspace.get(torch.device('cuda:0'), h.workspace_bytes(B, False))
    buf = torch.empty(B * 20680, device='cuda')
    for pitch, want in ((20671, 1), (20674, 1), (20000, 1), (20672, 0), (20680, 0)):
        rc = L.prk_smpl_forward(h.handle, _runtime.ptr(pose.cuda()), None, None, -1, B, _runtime.ptr(buf), pitch, _runtime.ptr(jd),
                                ws, ws_bytes, _runtime.stream_ptr(torch.device('cuda:0')))
        assert rc == want, (pitch, rc)
    torch.cuda.synchronize()
    m = synthetic_smpl('neutral')
    w = np.zeros((6890, 24), np.float32)
    w[:, :6] = 1.0 / 6
    lay6 = SMPL_Layer(model_data=SMPLModelData(**{**m.__dict__, 'weights': w}))
    h6 = lay6._handle(torch.device('cuda:0'))
    rc = L.prk_smpl_forward(h6.handle, _runtime.ptr(pose.cuda()), None, None, -1, B, _runtime.ptr(buf), 20672, _runtime.ptr(jd),
                            ws, ws_bytes, _runtime.stream_ptr(torch.device('cuda:0')))
    assert rc == 4                                                # PRK_ERR_UNSUPPORTED
    v6, _ = SMPL_Layer(model_data=lay6.smpl_data, aligned_verts=True)(pose.cuda())     # the layer falls back to a dense tensor
    assert v6.is_contiguous()


@pytest.mark.parametrize('pd_scale,sd_scale,vt_scale', [(1.0, 1.0, 1.0), (1e-4, 1.0, 1.0), (30.0, 0.1, 1.0), (1.0, 40.0, 3.0),
                                                           (0.0, 1.0, 1.0), (1e-7, 1e-7, 1.9)])
def test_blend_operand_scale_follows_the_model(pd_scale, sd_scale, vt_scale):
    """The blend runs as fp16 + e4m3 tensor-core products on operands multiplied by 2^S (csrc/prk_internal.h "K12 operand
    layout"); S is derived from max|posedirs|, max|shapedirs| and max|v_template| so that no part leaves the fp16 / e4m3
    range.  Models whose blend shapes are far smaller / larger than SMPL's, large betas and large rotations stay inside the
    1e-5 bound of north_star against the float64 oracle."""
    from poserisk_release_b200 import SMPL_Layer
    from poserisk_release_b200.model_provider import SMPLModelData
    m = synthetic_smpl('neutral')
    md = SMPLModelData((m.v_template * vt_scale).astype(np.float32), (m.shapedirs * sd_scale).astype(np.float32),
                       (m.posedirs * pd_scale).astype(np.float32), m.J_regressor, m.weights, m.betas, m.faces,
                       m.kintree_table, 'neutral', True)
    lay = SMPL_Layer(gender='neutral', model_root='unused', model_data=md)
    g = torch.Generator().manual_seed(21)
    for pose_s, beta_s in ((0.35, 1.0), (1.2, 3.0), (0.6, 10.0)):
        pose = torch.randn(300, 72, generator=g) * pose_s
        betas = torch.randn(300, 10, generator=g) * beta_s
        v, j = lay(pose.cuda(), betas.cuda())
        v_ref, j_ref = oracle.smpl_forward(md, pose.numpy(), betas.numpy())
        ev, ej = relerr(v.cpu().numpy(), v_ref), relerr(j.cpu().numpy(), j_ref)
        # error model: the e4m3 cross terms leave 2^-15.5 of the POSE blend displacement; 2e-6 of the vertex range for
        # SMPL-like magnitudes (displacements of centimetres), the north_star bound for blend shapes 30 x larger
        bound = 2e-6 if pd_scale <= 1.0 else TOL
        print(f'\n[pd x{pd_scale} sd x{sd_scale} vt x{vt_scale} pose x{pose_s} betas x{beta_s}] verts {ev:.2e} joints {ej:.2e}')
        assert ev < bound and ej < 2e-6, (pd_scale, sd_scale, vt_scale, pose_s, beta_s, ev, ej)


def test_fast_screening_gives_the_exact_records(engine):
    """Large float32 batches are scored with a float32 screening pass; frames with an angle within 1e-3 degrees of a ladder
    threshold (and tiny / large rotations, sy < 0.02) are re-evaluated in float64 (csrc/prk_score.cu euler_screen).  The
    records must be bit-identical to the all-float64 evaluation: 1M frames whose twelve scored joints are built from Euler
    angles sitting ON the thresholds up to offsets of 1e-7 .. 1e-2 degrees (the float32 rounding of the axis-angle vector
    then scatters them around the threshold at the 1e-5 degree level), gimbal-lock pitches and exact zeros included.  The
    reference evaluation is the small-batch kernel (float64 throughout, chunks of 16,384 frames), itself checked against
    the oracle on a sample here and in test_pose_to_scores_vs_oracle."""
    from scipy.spatial.transform import Rotation
    from poserisk_release_b200 import _runtime
    n = 1 << 20
    rng = np.random.default_rng(5)
    thr = np.array([0, 1, -1, 5, -5, 10, -10, 15, -15, 20, -20, 30, -30, 45, -45, 60, -60, 70, -70, 90, -90, 100, -100,
                    110, -110, 180, -180, 35, 125], np.float64)
    off = np.array([0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 5e-4, -5e-4, 9e-4, -9e-4, 1.1e-3, -1.1e-3,
                    2e-3, -2e-3, 1e-2, -1e-2], np.float64)
    pose = (rng.standard_normal((n, 24, 3)) * 0.5).astype(np.float32)
    scored = [3, 4, 5, 12, 13, 14, 16, 17, 18, 19, 20, 21]
    e = thr[rng.integers(0, len(thr), (n, 12, 3))] + off[rng.integers(0, len(off), (n, 12, 3))]
    rnd = rng.random((n, 12, 3)) < 0.5                     # half of the components: anything
    e = np.where(rnd, rng.uniform(-180, 180, (n, 12, 3)), e)
    e[..., 1] = np.clip(e[..., 1], -90, 90)               # pitch; +-90 = gimbal lock (sy ~ 0)
    rv = Rotation.from_euler('xyz', e.reshape(-1, 3), degrees=True).as_rotvec().reshape(n, 12, 3)
    pose[:, scored] = rv.astype(np.float32)
    pose[rng.random(n) < 0.05, 20] = 0.0                   # exact zero rotations
    pose[rng.random(n) < 0.01, 16] *= 1e-5                 # tiny rotations
    pose[rng.random(n) < 0.01, 17] *= 3.0                  # rotations beyond 4 rad
    pose = torch.from_numpy(pose.reshape(n, 72)).cuda()
    fast, _ = engine.euler_debug(pose, [], EXAMPLE_INFO)                       # one call: the large-batch kernel
    exact = torch.cat([engine.euler_debug(pose[i:i + 16384], [], EXAMPLE_INFO)[0] for i in range(0, n, 16384)])
    torch.cuda.synchronize()
    diff = (fast != exact).any(dim=1)
    assert int(diff.sum()) == 0, (int(diff.sum()), torch.nonzero(diff)[:5].tolist())
    # with debug joints (float64 rows for them, screening for the rest)
    ids = [12, 16, 17, 3]
    fast_d, eul = engine.euler_debug(pose[:300000], ids, EXAMPLE_INFO)
    assert torch.equal(fast_d, exact[:300000])
    from poserisk_release_b200 import axis_angle_to_euler_angle
    e_ref = axis_angle_to_euler_angle(pose[:300000].reshape(-1, 24, 3)[:, ids].reshape(-1, 3).cpu().numpy())
    assert np.array_equal(eul.cpu().numpy().reshape(-1, 3), np.asarray(e_ref).reshape(-1, 3))
    # and the oracle on a sample
    ref = oracle.score_pose(pose[:20000].cpu().numpy(), EXAMPLE_INFO)
    assert same_records(_runtime.records_to_numpy(fast[:20000]), ref).all()
