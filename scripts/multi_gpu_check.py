"""N-rank correctness of the sharded path: every rank scores its contiguous frame shard, the
32-byte records are all-gathered over NCCL, and the result must equal the 1-rank result."""
import os; os.environ.setdefault("PRK_SYNTHETIC_SMPL", "1")
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poserisk_release_b200.pipeline import PoseRiskEngine  # noqa: E402
from poserisk_release_b200.distributed import run_sharded, shard_range  # noqa: E402
from poserisk_release_b200 import _runtime  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
info = {"REBA": {"Legs_bilateral_weight_bearing/walking": 1, "Sitting": 1, "Load/Force Score": 0,
                 "Arm_supported_leaning_L": 0, "Arm_supported_leaning_R": 0, "Coupling": 0, "Activity_Score": 0},
        "RULA": {"Arm_supported_leaning_L": 0, "Arm_supported_leaning_R": 0, "A_Muscle_use_L": 0, "A_Muscle_use_R": 0,
                 "A_Load/Force_L": 0, "A_Load/Force_R": 0, "Legs_bilateral_weight_bearing": 0, "B_Muscle_use": 0,
                 "B_Load/Force": 0}}
n = 100003   # ragged shards
g = torch.Generator().manual_seed(0)
pose = (torch.randn(n, 72, generator=g) * 0.5).to(dev)
betas = torch.randn(n, 10, generator=g).to(dev)
eng = PoseRiskEngine(dev)
out = run_sharded(eng, pose, betas, None, info, want_verts=False)
single = eng.run(pose, betas, None, add_info=info, want_verts=False)
torch.cuda.synchronize()
assert out['scores'].shape == (n, 32)
assert torch.equal(out['scores'], single['scores']), 'sharded scores differ from single-rank scores'
lo, hi = shard_range(n, rank, world)
assert torch.equal(out['joints'], single['joints'][lo:hi])
# full-mesh shard
m = 3000
o2 = run_sharded(eng, pose[:m], betas[:m], None, info, want_verts=True)
s2 = eng.run(pose[:m], betas[:m], None, add_info=info, want_verts=True)
lo, hi = shard_range(m, rank, world)
assert torch.equal(o2['verts'], s2['verts'][lo:hi]) and torch.equal(o2['scores'], s2['scores'])
dist.barrier()
if rank == 0:
    print(f'multi-gpu check ok: world={world} frames={n}')
dist.destroy_process_group()
