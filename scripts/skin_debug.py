import ctypes as C, os, sys, subprocess, torch
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
buf = (C.c_ulonglong * 16)()
L.prk_skin_debug_read(buf, 1)
n = 5
for _ in range(n): eng.run(pose, betas, None, add_info=info)
torch.cuda.synchronize()
L.prk_skin_debug_read(buf, 0)
names2 = {9: 'commit -> epilogue warp0 wake (sum)', 10: 'last t_empty arrive -> mma wake (sum)', 11: 'builder wake -> mma wake (sum)', 12: 'count (epi warp0 items)', 13: 'commit(i) -> builder wake for i+2 (sum)'}
names = ['epi wait t_full (16 warps)', 'epi wait vp_full (16 warps)', 'builder wait w_empty (4 warps)', 'mma wait w_full', 'mma wait t_empty', 'tma wait vp_empty', '', '', 'kernel cycles (per CTA)']
ctas = 148
for k, nm in enumerate(names):
    if nm: print(f'{nm:36s} {buf[k] / n / ctas:12.0f} cycles per CTA per launch')

cnt = buf[12] / n
for k, nm in names2.items():
    print(f'{nm:44s} {buf[k] / n / max(cnt,1):10.1f} cycles per item   (raw {buf[k]})')
