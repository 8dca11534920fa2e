# Round-2 single-GPU validation: all GPU tests, smoke, bench with configs 3/4/5, joints-only timing, sanitizer attempt.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -30 > gpurun_out/pytest_r2.log
tail -5 gpurun_out/pytest_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r2.log
timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench_r2.err > gpurun_out/bench_r2.log
tail -c 600 gpurun_out/bench_r2.err
python scripts/show_bench.py < gpurun_out/bench_r2.log
# compute-sanitizer: does it run at all on this pool?  (small batch, racecheck + memcheck of the three kernels)
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
os.environ['PRK_SYNTHETIC_SMPL'] = '1'
import torch
from poserisk_release_b200 import PoseRiskEngine
eng = PoseRiskEngine('cuda:0')
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1])
out = eng.run((torch.randn(B, 72, generator=g) * 0.4).cuda(), torch.randn(B, 10, generator=g).cuda(), None,
              add_info={"REBA": {k: 0 for k in __import__('poserisk_release_b200')._lib.REBA_KEYS}, "RULA": {k: 0 for k in __import__('poserisk_release_b200')._lib.RULA_KEYS}})
torch.cuda.synchronize()
print('ok', float(out['verts'].abs().sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --kernel-regex kns=prk --launch-timeout 60 python /tmp/san.py 300 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
