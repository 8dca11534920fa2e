# quick A/B timing of the kernels at (mostly) un-capped clocks: short runs, no preload
for r in 1 2 3; do
  PRK_BENCH_PRELOAD_S=0 timeout 120 python bench.py --steps 20 --warmup 3 2>/dev/null | python scripts/show_bench.py | head -1
  sleep 2
done
