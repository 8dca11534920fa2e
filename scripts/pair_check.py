"""CTA-pair (cta_group::2) vertex kernel against the single-CTA kernel, and bulk-tensor-store output (16-byte aligned rows)
against the dense reference layout, frame by frame, for ragged batch sizes: every batch is run (a) with aligned rows, (b) dense,
(c) dense in 128-frame slices (one frame tile -> always the single-CTA kernel), and the vertices are compared bit for bit."""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1"); os.environ["PRK_PAIR"] = "2"
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poserisk_release_b200 import PoseRiskEngine, _lib
eng = PoseRiskEngine("cuda:0")
info = {"REBA": {k: 0 for k in _lib.REBA_KEYS}, "RULA": {k: 0 for k in _lib.RULA_KEYS}}
bad_total = 0
for B in (129, 256, 300, 384, 641, 1024, 1280, 1408, 1500, 1536, 1664, 2048, 3000, 4096, 4173, 9473):
    g = torch.Generator().manual_seed(B)
    pose = (torch.randn(B, 72, generator=g) * 0.6).cuda(); betas = torch.randn(B, 10, generator=g).cuda()
    v = eng.run(pose, betas, None, add_info=info)["verts"]                       # engine-allocated: aligned rows, TMA stores
    assert not v.is_contiguous() or B == 1
    dense = torch.full((B, 6890, 3), float("nan"), device="cuda")
    eng.run(pose, betas, None, add_info=info, verts_out=dense)
    ref = torch.cat([eng.run(pose[i:i + 128], betas[i:i + 128], None, add_info=info, verts_out=dense.new_empty((min(128, B - i), 6890, 3)))["verts"]
                     for i in range(0, B, 128)])
    torch.cuda.synchronize()
    bad = (v != ref).any(dim=2).any(dim=1) | (dense != ref).any(dim=2).any(dim=1)
    bad_total += int(bad.sum())
    print(B, "differing frames", int(bad.sum()), torch.nonzero(bad).flatten()[:6].tolist())
print("OK" if bad_total == 0 else "MISMATCH")
