"""Joints-only pose chain: lane-per-joint vs thread-per-frame kernel (PRK_CHAIN_WARP_MAX_JO = huge / 0), one process each."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get('CT_WORKER'):
    sys.path.insert(0, ROOT); os.environ['PRK_SYNTHETIC_SMPL'] = '1'
    import numpy as np, torch, bench
    from poserisk_release_b200 import PoseRiskEngine, _lib, _runtime
    dev = torch.device('cuda:0')
    eng = PoseRiskEngine(dev); L = _lib.lib(); h = eng.models['neutral']
    out = []
    for B in (2048, 4096, 8192, 16384, 32768, 65536, 131072):
        p, b, t = bench.counter_inputs(0, B, dev)
        j = torch.empty((B, 24, 3), device=dev)
        ws, ws_bytes, _keep = _runtime.workspace.get(dev, h.workspace_bytes(B, True))
        def run():
            _lib.check(L.prk_smpl_forward(h.handle, _runtime.ptr(p), _runtime.ptr(b), _runtime.ptr(t), -1, B, None, 0,
                                          _runtime.ptr(j), ws, ws_bytes, _runtime.stream_ptr(dev)))
        for _ in range(3): run()
        torch.cuda.synchronize()
        _lib.check(L.prk_profile_begin())
        for _ in range(10): run()
        torch.cuda.synchronize()
        ms = np.zeros(4); n = np.zeros(4, np.int64)
        _lib.check(L.prk_profile_end(ms.ctypes.data, n.ctypes.data))
        out.append('%d: %.1f us' % (B, ms[0] / n[0] * 1e3))
    print(' | '.join(out)); sys.exit(0)
for name, v in (('warp', str(1 << 40)), ('thread', '0')):
    r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, CT_WORKER='1', PRK_CHAIN_WARP_MAX_JO=v), capture_output=True, text=True)
    print(name, (r.stdout.strip().splitlines() or [r.stderr[-400:]])[-1], flush=True)
