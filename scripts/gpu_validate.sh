# Standard single-GPU validation on a B200 box: parity tests, smoke, bench (both arms).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest.log
cat gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 300 python bench.py --steps ${STEPS:-50} --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench.log | python scripts/show_bench.py
