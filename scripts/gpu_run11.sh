mkdir -p gpurun_out
export PRK_BENCH_PRELOAD_S=0
CMD="python bench.py --steps 3 --warmup 3"
$CMD > gpurun_out/plain17.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_blend -s 3 -c 1 -o gpurun_out/prof_fused_r1h $CMD > gpurun_out/ncu17.log 2>&1
tail -2 gpurun_out/ncu17.log
