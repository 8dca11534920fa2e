# ncu evidence for profiles/: launch list of one bench run and a full capture of the fused kernel.
mkdir -p gpurun_out
export PRK_BENCH_PRELOAD_S=0
CMD="python bench.py --steps 3 --warmup 3 --repeats 1 --skip-extra"
$CMD > gpurun_out/plain_prof.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain_prof2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_blend -s 6 -c 1 -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
# optional: full capture of the pose-chain kernel
if [ -n "$PROFILE_POSE" ]; then
ncu --set full --clock-control none --import-source on -k regex:pose_chain -s 6 -c 1 -o gpurun_out/prof_pose $CMD > gpurun_out/ncu_pose.log 2>&1
fi
# joints-only path (config 5 kernels): launch list of a 1M-frame call
cat > /tmp/jo.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
os.environ['PRK_SYNTHETIC_SMPL'] = '1'
import torch, bench
from poserisk_release_b200 import PoseRiskEngine
eng = PoseRiskEngine('cuda:0')
p, b, t = bench.counter_inputs(0, 1000000, 'cuda:0')
for _ in range(3):
    eng.run(p, b, t, add_info=bench.EXAMPLE_INFO, want_verts=False, debug_joints=bench.DEBUG_JOINTS)
torch.cuda.synchronize()
PY
python /tmp/jo.py > gpurun_out/plain_jo.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"pose_chain|score_pose" -c 6 --csv --log-file gpurun_out/launches_jo.csv python /tmp/jo.py > gpurun_out/ncu_jo.log 2>&1
tail -3 gpurun_out/launches_jo.csv | cut -c1-300
