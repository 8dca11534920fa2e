import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from poserisk_release_b200 import _runtime
from poserisk_release_b200.pipeline import PoseRiskEngine
dev = torch.device('cuda', 0)
B = 4096
eng = PoseRiskEngine(dev)
dev_in = [tuple(t.to(dev) for t in bench.make_inputs(i, B)) for i in range(8)]
host_in = [tuple(t.pin_memory() for t in bench.make_inputs(i, B)) for i in range(8)]
info_dev = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, dev)
verts = torch.empty((B, 6890, 3), device=dev)
dj = torch.empty((B, 24, 3), device=dev); ds = torch.empty((B, 32), dtype=torch.uint8, device=dev)
hj = torch.empty((B, 24, 3)).pin_memory(); hs = torch.empty((B, 32), dtype=torch.uint8).pin_memory()
def sd(i):
    p, b, t = dev_in[i % 8]; eng.run(p, b, t, add_info=info_dev, verts_out=verts, joints_out=dj, scores_out=ds)
def sh(i):
    p, b, t = host_in[i % 8]; eng.run_host(p, b, t, bench.EXAMPLE_INFO, None, hj, hs, verts_out=verts)
for name, fn in (('device', sd), ('host', sh)):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(200): fn(i)
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    print(f'{name}: enqueue {1e6*(t1-t0)/200:.1f} us/step, gpu {1e3*e0.elapsed_time(e1)/200:.1f} us/step')

import ctypes as C
from poserisk_release_b200 import _lib
def variant(name, joints, betas_trans):
    def fn(i):
        p, b, t = host_in[i % 8]
        eng.run_host(p, b if betas_trans else None, t if betas_trans else None, bench.EXAMPLE_INFO, None, hj if joints else None, hs, verts_out=verts)
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): fn(i)
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: gpu {1e3*e0.elapsed_time(e1)/200:.1f} us/step')
variant('host, no joints download', False, True)
variant('host, no betas/trans upload', True, False)
variant('host, neither', False, False)
for name, fn in (('device again', sd), ('host again', sh)):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): fn(i)
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: gpu {1e3*e0.elapsed_time(e1)/200:.1f} us/step')
