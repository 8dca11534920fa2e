mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rA 2>&1 | tail -120 > gpurun_out/pytest.log
echo "pytest exit: $?" >> gpurun_out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit: $?" >> gpurun_out/bench.log
tail -40 gpurun_out/pytest.log; tail -5 gpurun_out/bench.log
