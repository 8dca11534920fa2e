"""Does bracketing kernels with CUDA events change the step time?  50-step regions of the config-2 device path with no events,
events around every kernel, events around the fused kernel only -- interleaved, after a 1.5 s preload (sustained clocks)."""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from poserisk_release_b200 import PoseRiskEngine, _lib, _runtime
dev = torch.device("cuda:0"); eng = PoseRiskEngine(dev); L = _lib.lib()
B, K = 4096, 50
ins = [tuple(t.to(dev) for t in bench.make_inputs(i, B)) for i in range(8)]
info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, dev)
v = _runtime.aligned_verts(B, dev); j = torch.empty((B, 24, 3), device=dev); s = torch.empty((B, 32), dtype=torch.uint8, device=dev)
def step(i):
    p, b, t = ins[i % 8]
    eng.run(p, b, t, add_info=info, verts_out=v, joints_out=j, scores_out=s)
t_end = time.perf_counter() + 1.5
i = 0
while time.perf_counter() < t_end:
    step(i); i += 1
    if i % 64 == 0: torch.cuda.synchronize()
def region(mask):
    torch.cuda.synchronize()
    if mask is not None: _lib.check(L.prk_profile_begin_stages(mask))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    fused = None
    if mask is not None:
        ms = np.zeros(4); n = np.zeros(4, np.int64)
        _lib.check(L.prk_profile_end(ms.ctypes.data, n.ctypes.data))
        fused = ms[1] / max(n[1], 1) * 1e3
    return e0.elapsed_time(e1) / K * 1e3, fused
for rnd in range(4):
    for name, mask in (("none", None), ("all", 0xF), ("fused only", 2), ("pose only", 1), ("score only", 8), ("none", None)):
        st, f = region(mask)
        print(f"round {rnd} {name:10s} step {st:6.1f} us" + (f"  fused {f:6.1f} us" if f else ""), flush=True)
