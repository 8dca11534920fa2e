mkdir -p gpurun_out
( for al in 1 0; do echo "aligned $al"; AB_ALIGNED=$al AB_STEPS=10000 python scripts/fused_ab.py base bf16 base bf16; done
  echo "aligned 0 short"; AB_ALIGNED=0 python scripts/fused_ab.py base bf16 ) 2>&1 | tee gpurun_out/ab_r2_32.log
