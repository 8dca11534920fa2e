mkdir -p gpurun_out
( for x in 0 16 32 48 0 24 40 64; do echo "switch_cost16 $x"; PRK_SWITCH_COST16=$x AB_ALIGNED=1 python scripts/fused_ab.py base; done
  for x in 0 32 0 32; do echo "long switch_cost16 $x"; PRK_SWITCH_COST16=$x AB_ALIGNED=1 AB_STEPS=5000 python scripts/fused_ab.py base; done ) 2>&1 | tee gpurun_out/ab_r2_33.log
