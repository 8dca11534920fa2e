mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3 | tee gpurun_out/pytest_r2m.log
python scripts/pair_check.py 2>&1 | tail -1
timeout 600 python scripts/soak.py 300 11 2>&1 | tail -1 | tee gpurun_out/soak_r2c.log
