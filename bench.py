#!/usr/bin/env python
"""Headline benchmark: REBA+RULA scored frames/sec (BASELINE.json) on N GPUs of one box.

A step = one pass of the hot path (SMPL forward with full 6890-vertex mesh + joints +
Euler angles + REBA + RULA) over one batch of 4096 synthetic frames per GPU
(BASELINE.json configs[1]); at N>1 every rank processes its own batches (weak scaling) and
the per-frame score records are all-gathered with NCCL inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]          product (CUDA) arm
  python bench.py --impl reference [...]                        CPU arm: the oracle port of
        the reference path on all host threads (the reference itself is pure Python and
        cannot travel to the GPU box)

Prints ONE JSON line (rank 0).  `value` = device-resident inputs, CUDA-event timed;
`e2e` = the same metric through prk_pipeline_host with pinned HOST buffers (host->device
copies of pose/betas/trans and device->host copies of scores+joints inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FRAMES_PER_STEP = 4096
METRIC = "REBA+RULA scored frames/sec"
UNIT = "frames/s"
EXAMPLE_INFO = {"REBA": {"Legs_bilateral_weight_bearing/walking": 1, "Sitting": 1, "Load/Force Score": 0,
                         "Arm_supported_leaning_L": 0, "Arm_supported_leaning_R": 0, "Coupling": 0,
                         "Activity_Score": 0},
                "RULA": {"Arm_supported_leaning_L": 0, "Arm_supported_leaning_R": 0, "A_Muscle_use_L": 0,
                         "A_Muscle_use_R": 0, "A_Load/Force_L": 0, "A_Load/Force_R": 0,
                         "Legs_bilateral_weight_bearing": 0, "B_Muscle_use": 0, "B_Load/Force": 0}}

# algorithmic work per frame
FUSED_BYTES_PER_FRAME = 340 + 82680 + 288 + 32          # whole pipeline with ideal fusion: 83,340 B
GEMM_FLOP_PER_FRAME = 2 * (10 + 207) * 20670            # 8,970,780 algorithmic blend flop
GEMM_EXEC_FLOP_PER_FRAME = 216 * 44 * 2 * 96 * 16       # executed MMA work: 216 tiles x 44 MMAs (N=96, K=16) = 29.2 MFLOP


def config_dict(n_gpus):
    return {"workload": "BASELINE.json configs[1]: 4096-frame batch of random SMPL pose/betas/trans per GPU per step, "
                        "full mesh (6890 verts) + joints + REBA/RULA scores, synthetic SMPL-shaped neutral model",
            "frames_per_step_per_gpu": FRAMES_PER_STEP, "parallelism": f"frames sharded x{n_gpus}, replicated model"
            + (", NCCL all-gather of 32 B/frame score records each step" if n_gpus > 1 else ""),
            "l2": "8 distinct input batches in rotation; every step writes 339 MB of vertices (> 126 MB L2), "
                  "so no input or output line survives in L2 between steps; the 21 MB bf16 blend matrix is "
                  "meant to stay L2-resident"}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': float(d['hbm_gbs']), 'bf16_tflops': float(d['bf16_tflops']),
                'bf16_tflops_sustained': float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


def make_inputs(seed, B):
    import torch
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(B, 72, generator=g) * 0.35
    betas = torch.randn(B, 10, generator=g)
    trans = torch.randn(B, 3, generator=g) * 0.1
    return pose, betas, trans


# ----------------------------------------------------------------------------- CPU arm
def cpu_inputs(frames):
    from poserisk_release_b200.model_provider import synthetic_smpl
    return (synthetic_smpl('neutral'),) + tuple(x.numpy() for x in make_inputs(1234, frames))


def cpu_step(inputs):
    """One pass of the oracle port of the reference path (SMPL forward + Euler + REBA + RULA), all host threads."""
    from oracle import oracle
    m, pose, betas, trans = inputs
    oracle.smpl_forward(m, pose, betas, trans)
    oracle.score_pose(pose, EXAMPLE_INFO)


# every angle threshold of the REBA/RULA ladders, degrees
THRESHOLDS = np.array([0, 1, -1, 5, -5, 10, -10, 15, -15, 20, -20, 30, 45, -45, 60, -60, 70, -70, 90, -90,
                       100, -100, 110, -110], dtype=np.float64)
SCORED_JOINTS = [3, 4, 5, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def parity_report(eng, frames, dev):
    """GPU results against the oracle on the CPU leg's own sample (last row): vertex / joint error,
    Euler error, exact-match rate of the scores, frames with a scored angle within 1e-3 degrees of a ladder threshold
    and how many of those differ.  Part of the cpu_baseline leg: the oracle is the checker here, never the product."""
    import torch
    from oracle import oracle
    from poserisk_release_b200 import _runtime
    m, pose, betas, trans = cpu_inputs(frames)
    v_ref, j_ref = oracle.smpl_forward(m, pose, betas, trans)
    rec_ref, eul_ref = oracle.score_pose(pose, EXAMPLE_INFO, want_euler=True)
    tp = torch.from_numpy(pose).to(dev)
    out = eng.run(tp, torch.from_numpy(betas).to(dev), torch.from_numpy(trans).to(dev), add_info=EXAMPLE_INFO)
    _, eul = eng.euler_debug(tp, list(range(24)), EXAMPLE_INFO)
    torch.cuda.synchronize(dev)
    v = out['verts'].cpu().numpy(); j = out['joints'].cpu().numpy()
    rec = _runtime.records_to_numpy(out['scores'])
    eul = eul.cpu().numpy().reshape(frames, 24, 3)
    eul_ref = np.asarray(eul_ref).reshape(frames, 24, 3)
    same = ((rec['reba_score'] == rec_ref['reba_score']) & (rec['rula_score'] == rec_ref['rula_score'])
            & (rec['reba_parts'] == rec_ref['reba_parts']).all(axis=1) & (rec['rula_parts'] == rec_ref['rula_parts']).all(axis=1))
    used = eul_ref[:, SCORED_JOINTS, :].reshape(frames, -1)
    near = (np.abs(used[:, :, None] - THRESHOLDS[None, None, :]).min(axis=2) < 1e-3).any(axis=1)
    return {"frames": int(frames), "against": "oracle/poserisk_oracle.c on the same inputs",
            "verts_max_abs_over_max_abs_ref": float(np.abs(v - v_ref).max() / np.abs(v_ref).max()),
            "verts_frobenius_rel": float(np.linalg.norm((v - v_ref).ravel()) / np.linalg.norm(v_ref.ravel())),
            "joints_max_abs_over_max_abs_ref": float(np.abs(j - j_ref).max() / np.abs(j_ref).max()),
            "euler_max_abs_err_deg": float(np.abs(eul - eul_ref).max()),
            "scores_exact_match_rate": float(same.mean()),
            "frames_near_threshold_1e-3deg": int(near.sum()), "of_those_differing": int((near & ~same).sum())}


def cpu_throughput(frames, repeats=1):
    from oracle import oracle
    inputs = cpu_inputs(frames)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_step(inputs)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return frames / best, best, oracle.num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    oracle.use_all_cores()
    # one step = the GPU arm's batch (4096 frames), shrunk only if this host needs more than ~2 s for it
    fps0, _, threads = cpu_throughput(512, repeats=2)
    frames = int(min(FRAMES_PER_STEP, max(256, fps0 * 2.0)))
    inputs = cpu_inputs(frames)                      # synthetic inputs are made once, outside the timed steps
    for _ in range(args.warmup):
        cpu_step(inputs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(inputs)
    dt = time.perf_counter() - t0
    value = frames * args.steps / dt
    sample = f"{frames} frames per step (same distribution as the GPU arm), {args.steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.power = []
        self._stop_evt = threading.Event()
        self.ok = False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for cand in ('GPU-' + uuid, uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            while not self._stop_evt.is_set():
                self.samples.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.02)
        except Exception:
            self.ok = False

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    if os.environ.get('PRK_BENCH_DEBUG'):      # stacks of every thread after N seconds (hang diagnosis)
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['PRK_BENCH_DEBUG']), exit=True)
    import torch
    import torch.distributed as dist
    from poserisk_release_b200 import _lib, _runtime
    from poserisk_release_b200.pipeline import PoseRiskEngine
    from poserisk_release_b200.distributed import all_gather_rows

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    B = FRAMES_PER_STEP
    K, W = args.steps, args.warmup
    eng = PoseRiskEngine(dev)
    L = _lib.lib()
    n_rot = 8
    dev_in = [tuple(t.to(dev) for t in make_inputs(1000 * rank + i, B)) for i in range(n_rot)]
    host_in = [tuple(t.pin_memory() for t in make_inputs(1000 * rank + i, B)) for i in range(n_rot)]
    info_dev = _runtime.addinfo_tensor(EXAMPLE_INFO, dev)
    verts = torch.empty((B, 6890, 3), dtype=torch.float32, device=dev)
    d_joints = torch.empty((B, 24, 3), dtype=torch.float32, device=dev)
    d_scores = torch.empty((B, 32), dtype=torch.uint8, device=dev)
    h_joints = torch.empty((B, 24, 3), dtype=torch.float32).pin_memory()
    h_scores = torch.empty((B, 32), dtype=torch.uint8).pin_memory()

    # N > 1: every step all-gathers its 32-byte score records on the compute stream, inside the timed region
    def gather_scores(scores_dev):
        all_gather_rows(scores_dev, B * world)

    def step_device(i):
        p, b, t = dev_in[i % n_rot]
        out = eng.run(p, b, t, add_info=info_dev, verts_out=verts, joints_out=d_joints, scores_out=d_scores)
        if world > 1:
            gather_scores(out['scores'])
        return out

    def step_host(i):
        p, b, t = host_in[i % n_rot]
        eng.run_host(p, b, t, EXAMPLE_INFO, None, h_joints, h_scores, verts_out=verts)
        if world > 1:
            gather_scores(eng.host_scores_device)

    def barrier():
        if world > 1:
            torch.cuda.synchronize(dev)
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # preload: keep the GPU busy ~1.5 s so clocks settle and the sampler sees load (not warm-up steps).
    # Time-bounded, so every rank runs its own number of iterations: NO collective in here.
    t_end = time.perf_counter() + float(os.environ.get('PRK_BENCH_PRELOAD_S', '1.5'))
    i = 0
    while time.perf_counter() < t_end:
        p, b, t = dev_in[i % n_rot]
        eng.run(p, b, t, add_info=info_dev, verts_out=verts, joints_out=d_joints, scores_out=d_scores)
        i += 1
        if i % 64 == 0:
            torch.cuda.synchronize(dev)
    for i in range(W):
        step_device(i)
    launches0 = _lib.launch_count()
    ms = timed(step_device, K)                       # the timed region: K steps, nothing else on the streams
    launches = _lib.launch_count() - launches0
    # second pass of K steps with a CUDA-event pair around every kernel (on the stream it is launched on):
    # per-kernel durations for the roofline.  Kept out of the timed region: the event records cost ~5 %.
    _lib.check(L.prk_profile_begin())
    ms_prof = timed(step_device, K)
    stage_ms = (np.zeros(4), np.zeros(4, np.int64))
    _lib.check(L.prk_profile_end(stage_ms[0].ctypes.data, stage_ms[1].ctypes.data))
    # e2e: host buffers through prk_pipeline_host
    for i in range(max(W, 3)):
        step_host(i)
    ms_e2e = timed(step_host, K)
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks = load_peaks()
        value = B * world * K / (ms * 1e-3)
        e2e_value = B * world * K / (ms_e2e * 1e-3)
        st_ms, st_n = stage_ms
        frames_timed = B * K
        per_stage = {}
        names = ('pose_chain', 'fused_blend_skin', None, 'scoring')
        for k, nm in enumerate(names):
            if nm:
                per_stage[nm] = {"ms_total": float(st_ms[k]), "launches": int(st_n[k]),
                                 "ms_per_launch": float(st_ms[k] / st_n[k]) if st_n[k] else None}
        per_stage["note"] = f"event pass: {ms_prof / K:.4f} ms/step with the per-kernel event pairs"
        fused_s = st_ms[1] * 1e-3
        hbm_gbs = FUSED_BYTES_PER_FRAME * frames_timed / fused_s / 1e9 if fused_s > 0 else 0.0
        gemm_tf = GEMM_FLOP_PER_FRAME * frames_timed / fused_s / 1e12 if fused_s > 0 else 0.0
        gemm_exec_tf = GEMM_EXEC_FLOP_PER_FRAME * frames_timed / fused_s / 1e12 if fused_s > 0 else 0.0
        tensor_peak = peaks['bf16_tflops_sustained']
        traffic = None      # dram bytes per launch of the fused kernel from the committed ncu capture
        tp = os.path.join(ROOT, 'profiles', 'fused_traffic.json')
        if os.path.isfile(tp):
            with open(tp) as f:
                traffic = json.load(f).get('dram_bytes_per_launch')
        dominant = {"kernel": "fused_blend_skin_kernel (tcgen05 blend GEMM + TMEM-resident skinning)", "bound": "hbm",
                    "achieved": hbm_gbs, "peak": peaks['hbm_gbs'], "unit": "GB/s", "frac": hbm_gbs / peaks['hbm_gbs'],
                    "traffic": traffic, "peak_source": peaks['source'],
                    "algorithmic_bytes_per_frame": FUSED_BYTES_PER_FRAME, "frames_per_launch": B}
        other = {"kernel": "fused_blend_skin_kernel, blend GEMM part", "bound": "tensor", "achieved": gemm_tf,
                 "peak": tensor_peak, "unit": "TFLOP/s", "frac": gemm_tf / tensor_peak, "traffic": None,
                 "executed_mma": {"achieved": gemm_exec_tf, "frac": gemm_exec_tf / tensor_peak,
                                  "note": "bf16 split precision (hi*hi + lo*hi + hi*lo, 3-way for betas) + N padding: "
                                          "29.2 MFLOP executed per 8.97 MFLOP algorithmic"},
                 "peak_source": peaks['source'] + " (sustained bf16, kernel timed inside a long step)"}
        cpu_baseline = parity = None
        if world == 1:   # bounded CPU sample of the same workload (rank 0 at N=1 only)
            from oracle import oracle
            oracle.use_all_cores()
            fps0, _, threads = cpu_throughput(256)
            n_cpu = int(min(8 * FRAMES_PER_STEP, max(512, fps0 * 12)))
            cpu_fps, cpu_dt, threads = cpu_throughput(n_cpu)
            parity = parity_report(eng, min(n_cpu, FRAMES_PER_STEP), dev)
            cpu_baseline = {"value": cpu_fps, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{n_cpu} frames of the same workload, {cpu_dt:.1f} s, C/OpenMP oracle "
                                      f"port of the reference path (oracle/poserisk_oracle.c)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": B * (72 + 10 + 3) * 4 + 64,
                        "d2h_bytes_per_step": B * (72 * 4 + 32),
                        "note": "vertices stay in HBM (the reference reads them only for a debug .obj)"},
                "gpu_launches": int(launches),
                "roofline": dominant, "roofline_other": other, "stages": per_stage,
                "cpu_baseline": cpu_baseline, "parity": parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
