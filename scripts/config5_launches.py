"""One rank's config-5 job (joints-only, 1M frames, scores + debug Euler rows, exchange) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('PRK_SYNTHETIC_SMPL', '1')
import torch, bench
from poserisk_release_b200 import PoseRiskEngine, _runtime
from poserisk_release_b200.distributed import ScoreExchange
dev = torch.device('cuda', 0)
eng = PoseRiskEngine(dev)
n = 1_000_000
pose, betas, trans = bench.counter_inputs(0, n, dev)
info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, dev)
joints = torch.empty((n, 24, 3), dtype=torch.float32, device=dev)
scores = torch.empty((n, 32), dtype=torch.uint8, device=dev)
euler = torch.empty((n, len(bench.DEBUG_JOINTS), 3), dtype=torch.float64, device=dev)
ex = ScoreExchange(n, dev, len(bench.DEBUG_JOINTS), None, 'auto')
def job():
    o = eng.run(pose, betas, trans, add_info=info, want_verts=False, joints_out=joints, scores_out=scores,
                debug_joints=bench.DEBUG_JOINTS, euler_out=euler, exchange=ex, frame_offset=0)
    return ex.collect(o['scores'], 0, o['euler'])
for _ in range(3): job()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): job()
e1.record(); torch.cuda.synchronize()
print('config-5 job: %.3f ms' % (e0.elapsed_time(e1) / 10))
