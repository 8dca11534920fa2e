# ncu evidence for profiles/: launch list of one bench run and a full capture of the fused kernel.
mkdir -p gpurun_out
export PRK_BENCH_PRELOAD_S=0
CMD="python bench.py --steps 3 --warmup 3"
$CMD > gpurun_out/plain_prof.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain_prof2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_blend -s 6 -c 1 -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
# optional: full capture of the pose-chain kernel
if [ -n "$PROFILE_POSE" ]; then
ncu --set full --clock-control none --import-source on -k regex:pose_chain -s 6 -c 1 -o gpurun_out/prof_pose $CMD > gpurun_out/ncu_pose.log 2>&1
fi
