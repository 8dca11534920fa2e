import os, sys
sys.path.insert(0, os.getcwd()); os.environ['PRK_SYNTHETIC_SMPL'] = '1'
import numpy as np, torch, bench
from poserisk_release_b200 import PoseRiskEngine, _lib, _runtime
eng = PoseRiskEngine('cuda:0'); L = _lib.lib()
for B in (8192, 16384, 32768, 65536):
    p, b, t = bench.counter_inputs(0, B, 'cuda:0')
    v = _runtime.aligned_verts(B, torch.device('cuda:0'))
    info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, torch.device('cuda:0'))
    for _ in range(2): eng.run(p, b, t, add_info=info, verts_out=v)
    torch.cuda.synchronize()
    _lib.check(L.prk_profile_begin())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.run(p, b, t, add_info=info, verts_out=v)
    e1.record(); torch.cuda.synchronize()
    ms = np.zeros(4); n = np.zeros(4, np.int64)
    _lib.check(L.prk_profile_end(ms.ctypes.data, n.ctypes.data))
    print(B, 'call %.3f ms' % (e0.elapsed_time(e1) / 3), 'pose %.1f us/launch x%d' % (ms[0] / n[0] * 1e3, n[0] // 3), 'fused %.1f us/launch' % (ms[1] / n[1] * 1e3), 'checksum %.6e' % float(v[:64].double().sum()))
