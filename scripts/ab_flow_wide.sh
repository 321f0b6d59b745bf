#!/bin/bash
# A/B: one witness over a 255-bit field, barrier kernels (ZKB_FLOW=0) vs the flag-word dataflow launch (k_levels_flow_wide)
for cfg in "bls381 20 0" "bls381 20 4096" "bls381 16 0" "bn254 18 1024"; do
  for f in 0 1; do ZKB_FLOW_WIDE=1 ZKB_FLOW=$f timeout 300 python scripts/flow_wide_once.py $cfg 2>&1 | tail -1; done
done
