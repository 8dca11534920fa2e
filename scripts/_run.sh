mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_f8.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -3 | tee gpurun_out/pytest_f8.log
( AB_ALIGNED=1 python scripts/fused_ab.py base m31 base m31
  AB_ALIGNED=1 AB_STEPS=2000 python scripts/fused_ab.py base m31 base m31 ) 2>&1 | tee gpurun_out/ab_r2_30.log
