"""Config 5 of BASELINE.json on one GPU: 1M frames, joints-only (no vertices), REBA+RULA scores and the
debug Euler sequences of four joints.  CUDA-event timing, inputs resident in HBM."""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from poserisk_release_b200 import _runtime
from poserisk_release_b200.pipeline import PoseRiskEngine
dev = torch.device('cuda', 0)
n = 1_000_000
eng = PoseRiskEngine(dev)
g = torch.Generator().manual_seed(0)
pose = (torch.randn(n, 72, generator=g) * 0.35).to(dev); betas = torch.randn(n, 10, generator=g).to(dev)
info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, dev)
joints = torch.empty((n, 24, 3), device=dev); scores = torch.empty((n, 32), dtype=torch.uint8, device=dev)
ids = [12, 16, 17, 3]
def run():
    eng.run(pose, betas, None, add_info=info, want_verts=False, joints_out=joints, scores_out=scores)
def run_dbg():
    eng.euler_debug(pose, ids, info)
for name, fn in (('joints + scores', run), ('scores + debug Euler of 4 joints', run_dbg)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{name}: {ms:.3f} ms per 1M frames = {n / ms / 1e3:.1f} M frames/s')

# per-kernel split (CUDA-event pairs around each launch)
import numpy as np
from poserisk_release_b200 import _lib
L = _lib.lib()
_lib.check(L.prk_profile_begin())
for _ in range(10): run()
torch.cuda.synchronize()
ms = np.zeros(4); cnt = np.zeros(4, np.int64)
_lib.check(L.prk_profile_end(ms.ctypes.data, cnt.ctypes.data))
print(f'pose chain (joints only): {ms[0] / 10:.3f} ms   scoring: {ms[3] / 10:.3f} ms  per 1M frames')
