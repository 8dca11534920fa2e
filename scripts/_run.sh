mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 500 -x 2>&1 | tail -2 | tee gpurun_out/pytest_r2n.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_r2c.log
timeout 400 python bench.py 2> gpurun_out/bench_r2l.err > gpurun_out/bench_r2l.log
python scripts/show_bench.py < gpurun_out/bench_r2l.log | cut -c1-200
export PRK_BENCH_PRELOAD_S=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --repeats 1 --skip-extra > gpurun_out/ncu_launch.log 2>&1
