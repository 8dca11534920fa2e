"""One rank's share of BASELINE.json config 3 on one GPU: 262,144 frames (1M frames / 4 GPUs) with the full mesh,
joints and REBA/RULA scores in ONE call (four 65,536-frame pose-chain / vertex-kernel pairs), device-resident inputs,
CUDA-event timing; the first 4096 frames are compared bit for bit with a 4096-frame call."""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from poserisk_release_b200 import _runtime
from poserisk_release_b200.pipeline import PoseRiskEngine
dev = torch.device('cuda', 0)
n = int(os.environ.get('FRAMES', 262144))
eng = PoseRiskEngine(dev)
g = torch.Generator().manual_seed(0)
pose = (torch.randn(n, 72, generator=g) * 0.35).to(dev)
betas = torch.randn(n, 10, generator=g).to(dev)
trans = (torch.randn(n, 3, generator=g) * 0.1).to(dev)
info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, dev)
verts = torch.empty((n, 6890, 3), device=dev)
joints = torch.empty((n, 24, 3), device=dev); scores = torch.empty((n, 32), dtype=torch.uint8, device=dev)
def run():
    eng.run(pose, betas, trans, add_info=info, verts_out=verts, joints_out=joints, scores_out=scores)
run(); torch.cuda.synchronize()
small = eng.run(pose[:4096], betas[:4096], trans[:4096], add_info=info)
torch.cuda.synchronize()
same = bool(torch.equal(small['verts'], verts[:4096]) and torch.equal(small['scores'], scores[:4096]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f'{n} frames, full mesh + joints + scores: {ms:.2f} ms = {n / ms / 1e3:.2f} M frames/s, '
      f'{n * 83340 / ms / 1e6:.0f} GB/s of algorithmic traffic; first 4096 frames identical to a 4096-frame call: {same}')
