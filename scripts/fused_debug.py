"""Phase timers of the fused kernel's epilogue warps (library built with PRK_FUSED_DEBUG=1)."""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poserisk_release_b200 import _lib
from poserisk_release_b200.pipeline import PoseRiskEngine
L = _lib.lib()
eng = PoseRiskEngine('cuda:0')
info = {"REBA": {k: 0 for k in _lib.REBA_KEYS}, "RULA": {k: 0 for k in _lib.RULA_KEYS}}
B = 4096
pose = torch.randn(B, 72, device='cuda') * 0.35; betas = torch.randn(B, 10, device='cuda')
for _ in range(3): eng.run(pose, betas, None, add_info=info)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 8)()
L.prk_fused_debug_read(buf, 1)
n = 5
for _ in range(n): eng.run(pose, betas, None, add_info=info)
torch.cuda.synchronize()
L.prk_fused_debug_read(buf, 0)
warps = buf[7]
names = ['wait wfull', 'wait tfull', 'gather issue -> data (per vertex, summed)', 'transpose + store (per half, summed)', 'epilogue warp total']
for k, nm in enumerate(names):
    print(f'{nm:44s} {buf[k] / max(warps, 1):12.0f} cycles per warp per launch')
