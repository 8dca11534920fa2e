echo warp; PRK_CHAIN_WARP_MAX=100000000 python scripts/chain_threshold.py 2>&1 | grep call
echo thread; PRK_CHAIN_WARP_MAX=0 python scripts/chain_threshold.py 2>&1 | grep call
