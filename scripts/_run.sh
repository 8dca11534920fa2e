mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 900 -s -k "blend_operand or fast_screening" 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/pytest_r2i.log
