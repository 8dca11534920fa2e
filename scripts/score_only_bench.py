import os, sys
sys.path.insert(0, os.getcwd())
os.environ['PRK_SYNTHETIC_SMPL'] = '1'
import torch, bench
from poserisk_release_b200 import PoseRiskEngine, _runtime
eng = PoseRiskEngine('cuda:0')
n = 1000000
g = torch.Generator().manual_seed(0)
pose = (torch.randn(n, 72, generator=g) * 0.35).cuda()
info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, torch.device('cuda:0'))
for ids in ([], [12, 16, 17, 3]):
    for _ in range(3): eng.euler_debug(pose, ids, info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.euler_debug(pose, ids, info)
    e1.record(); torch.cuda.synchronize()
    print('debug joints', len(ids), ': %.3f ms per 1M frames' % (e0.elapsed_time(e1) / 10), flush=True)
