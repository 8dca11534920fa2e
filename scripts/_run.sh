mkdir -p gpurun_out
python scripts/score_only_bench.py 2>&1 | grep "debug joints"
python scripts/config5_launches.py 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -3 | tee gpurun_out/pytest_r2l.log
