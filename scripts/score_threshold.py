"""Crossover between the 16-lanes-per-frame scoring kernel and the thread-per-frame kernel (float32 screening):
PRK_SCORE_LANES_MAX=0 forces the thread kernel, a huge value the lanes kernel.  One process per setting."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get('ST_WORKER'):
    sys.path.insert(0, ROOT)
    os.environ['PRK_SYNTHETIC_SMPL'] = '1'
    import torch, bench
    from poserisk_release_b200 import PoseRiskEngine, _runtime
    eng = PoseRiskEngine('cuda:0')
    info = _runtime.addinfo_tensor(bench.EXAMPLE_INFO, torch.device('cuda:0'))
    out = []
    for n in (2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144):
        g = torch.Generator().manual_seed(0)
        pose = (torch.randn(n, 72, generator=g) * 0.35).cuda()
        for _ in range(5): eng.euler_debug(pose, [], info)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): eng.euler_debug(pose, [], info)
        e1.record(); torch.cuda.synchronize()
        out.append('%d: %.1f us' % (n, e0.elapsed_time(e1) / 20 * 1e3))
    print(' | '.join(out))
    sys.exit(0)
for name, v in (('lanes', str(1 << 40)), ('thread', '0')):
    r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, ST_WORKER='1', PRK_SCORE_LANES_MAX=v),
                       capture_output=True, text=True)
    print(name, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
