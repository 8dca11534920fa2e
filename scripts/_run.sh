mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -4 | tee gpurun_out/pytest_r2g.log
( AB_ALIGNED=1 python scripts/fused_ab.py base m29 base m29
  AB_ALIGNED=1 AB_STEPS=2000 python scripts/fused_ab.py base m29 ) 2>&1 | tee gpurun_out/ab_r2_31.log
