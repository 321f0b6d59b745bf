#!/bin/bash
# A/B of the R1CS term stream: z limbs of the next 1 or 2 terms in registers x resident CTAs per SM (register cap)
for cfg in "1 3" "2 3" "2 2" "1 2"; do
  set -- $cfg
  ZKB_EXTRA_NVCC_FLAGS="-DZKB_R1CS_ZDEPTH=$1 -DZKB_R1CS_MIN_CTAS=$2" python -c "import __graft_entry__ as g; g.build(force=True)" >/dev/null 2>&1
  echo "zdepth=$1 min_ctas=$2 single: $(python scripts/r1cs_once.py 22 1 | tail -1 | cut -c1-120)"
  echo "zdepth=$1 min_ctas=$2 batch64: $(python scripts/r1cs_once.py 18 64 | tail -1 | cut -c1-120)"
done
