"""A/B timing of kernel variants on one GPU: python scripts/fused_ab.py name[:dbg] ... (libs from build/variants/).

Every variant runs in its own process (PRK_LIB), 4096 frames per step; prints the per-kernel CUDA-event times and a
checksum of the vertices so a variant that changes results is seen at once.
"""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from poserisk_release_b200 import _lib, _runtime
    from poserisk_release_b200.pipeline import PoseRiskEngine
    B = int(os.environ.get('AB_FRAMES', '4096'))
    steps = int(os.environ.get('AB_STEPS', '40'))
    dev = torch.device('cuda', 0)
    eng = PoseRiskEngine(dev)
    L = _lib.lib()
    info = {"REBA": {k: 0 for k in _lib.REBA_KEYS}, "RULA": {k: 0 for k in _lib.RULA_KEYS}}
    info_dev = _runtime.addinfo_tensor(info, dev)
    ins = []
    for i in range(4):
        g = torch.Generator().manual_seed(i)
        ins.append(((torch.randn(B, 72, generator=g) * 0.35).to(dev), torch.randn(B, 10, generator=g).to(dev),
                    (torch.randn(B, 3, generator=g) * 0.1).to(dev)))
    # AB_ALIGNED=1: vertex rows 16-byte aligned (pitch 20672 floats) -> bulk tensor stores; default: the dense reference layout
    verts = _runtime.aligned_verts(B, dev) if os.environ.get('AB_ALIGNED') == '1' else torch.empty((B, 6890, 3), dtype=torch.float32, device=dev)
    joints = torch.empty((B, 24, 3), dtype=torch.float32, device=dev)
    scores = torch.empty((B, 32), dtype=torch.uint8, device=dev)

    def run(i):
        p, b, t = ins[i % 4]
        eng.run(p, b, t, add_info=info_dev, verts_out=verts, joints_out=joints, scores_out=scores)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    best = None
    nv = None
    try:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    except Exception:
        pass
    clk = pw = 0
    for rep in range(3):
        _lib.check(L.prk_profile_begin())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            run(i)
        e1.record()
        if nv is not None and steps >= 1000:      # long runs: clocks / power while the queue is still draining
            import time
            time.sleep(0.05)
            clk = pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
            pw = pynvml.nvmlDeviceGetPowerUsage(nv) / 1000.0
        torch.cuda.synchronize()
        ms = np.zeros(4); n = np.zeros(4, np.int64)
        _lib.check(L.prk_profile_end(ms.ctypes.data, n.ctypes.data))
        us = ms / np.maximum(n, 1) * 1e3
        if best is None or us[1] < best[1]:
            best = us
        last = (us, e0.elapsed_time(e1) / steps * 1e3)
    plain = 1e9
    for rep in range(3):          # without the per-kernel event pairs (they sit between the kernels)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        plain = min(plain, e0.elapsed_time(e1) / steps * 1e3)
    run(0)
    torch.cuda.synchronize()
    v = verts.double()
    chk = float(v.sum().item()); chk2 = float((v * v).sum().item())
    print('fused best %.1f last %.1f us | pose %.1f score %.1f | step %.1f us (plain %.1f) | checksum %.9e %.9e | %d MHz %.0f W'
          % (best[1], last[0][1], last[0][0], last[0][3], last[1], plain, chk, chk2, clk, pw), flush=True)


if __name__ == '__main__':
    if os.environ.get('AB_WORKER'):
        worker()
        sys.exit(0)
    for spec in sys.argv[1:]:
        name, _, dbg = spec.partition(':')
        env = dict(os.environ, AB_WORKER='1', PRK_FUSED_DBG=dbg or '0')
        if name != 'base':
            env['PRK_LIB'] = os.path.join(ROOT, 'build', 'variants', f'lib_{name}.so')
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=120)
        out = (r.stdout.strip().splitlines() or ['(no output)'])[-1]
        print('%-16s %s%s' % (spec, out, '' if r.returncode == 0 else '  rc=%d %s' % (r.returncode, r.stderr[-400:])), flush=True)
