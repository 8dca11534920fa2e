"""One small forward through the vertex kernel, checked against 128-frame slices (debug helper): python scripts/one_forward.py B"""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poserisk_release_b200 import PoseRiskEngine, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
eng = PoseRiskEngine("cuda:0")
info = {"REBA": {k: 0 for k in _lib.REBA_KEYS}, "RULA": {k: 0 for k in _lib.RULA_KEYS}}
g = torch.Generator().manual_seed(B)
pose = (torch.randn(B, 72, generator=g) * 0.6).cuda(); betas = torch.randn(B, 10, generator=g).cuda()
v = torch.full((B, 6890, 3), float('nan'), device='cuda')
eng.run(pose, betas, None, add_info=info, verts_out=v)
torch.cuda.synchronize()
print('ran', B, 'nan frames', int(torch.isnan(v).any(dim=2).any(dim=1).sum()), 'checksum', float(v.double().sum()))
