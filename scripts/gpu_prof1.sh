# launch list + full ncu captures of the two dominant kernels (round 1, baseline kernels)
mkdir -p gpurun_out
export PRK_BENCH_PRELOAD_S=0
CMD="python bench.py --steps 5 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:skin_kernel -s 14 -c 2 -o gpurun_out/prof_skin_r1a $CMD > gpurun_out/ncu_skin.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blend_gemm_kernel -s 14 -c 2 -o gpurun_out/prof_gemm_r1a $CMD > gpurun_out/ncu_gemm.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"pose_chain_kernel|score_pose_kernel" -s 4 -c 2 -o gpurun_out/prof_small_r1a $CMD > gpurun_out/ncu_small.log 2>&1
tail -3 gpurun_out/ncu_*.log; ls -la gpurun_out
