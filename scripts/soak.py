"""Soak test of the vertex kernel (no compute-sanitizer on the GPU pool): random ragged batch sizes, every batch run three times
(aligned rows / dense / aligned again, NaN-filled outputs, a second stream generating unrelated traffic) and compared bit for
bit; every 10th batch is also compared with 128-frame slices (single-CTA kernel).  python scripts/soak.py [iterations] [seed]"""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poserisk_release_b200 import PoseRiskEngine, _lib, _runtime
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
eng = PoseRiskEngine(dev, genders=("neutral", "female"))
info = {"REBA": {k: 0 for k in _lib.REBA_KEYS}, "RULA": {k: 0 for k in _lib.RULA_KEYS}}
side = torch.cuda.Stream()
noise = torch.empty(32 << 20, device=dev)
bad = 0
for it in range(n_iter):
    B = int(rng.choice([rng.integers(1, 300), rng.integers(300, 3000), rng.integers(3000, 12000)]))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    pose = (torch.randn(B, 72, generator=g) * 0.5).to(dev); betas = torch.randn(B, 10, generator=g).to(dev)
    trans = (torch.randn(B, 3, generator=g) * 0.1).to(dev)
    gender = ("neutral", "female")[it % 2]
    outs = []
    for k in range(3):
        if k == 1:
            v = torch.full((B, 6890, 3), float("nan"), device=dev)
        else:
            v = _runtime.aligned_verts(B, dev); v.fill_(float("nan"))
        with torch.cuda.stream(side):
            noise.normal_()
        r = eng.run(pose, betas, trans, add_info=info, gender=gender, verts_out=v)
        outs.append((v, r["joints"].clone(), r["scores"].clone()))
    torch.cuda.synchronize()
    ok = all(torch.equal(outs[0][j], outs[k][j]) for k in (1, 2) for j in range(3)) and not torch.isnan(outs[0][0]).any()
    if ok and it % 10 == 0:
        ref = torch.cat([eng.run(pose[i:i + 128], betas[i:i + 128], trans[i:i + 128], add_info=info, gender=gender,
                                 verts_out=torch.empty((min(128, B - i), 6890, 3), device=dev))["verts"] for i in range(0, B, 128)])
        torch.cuda.synchronize()
        ok = torch.equal(ref, outs[1][0])
    if not ok:
        bad += 1
        print("MISMATCH at iteration", it, "B", B, flush=True)
    del outs
print(f"{n_iter} batches, {bad} mismatches", flush=True)
sys.exit(1 if bad else 0)
