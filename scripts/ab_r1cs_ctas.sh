#!/bin/bash
# A/B of the R1CS check kernel's resident CTAs per SM (3: 80 registers, 4: 64 registers + spills)
for m in 3 4; do
  ZKB_EXTRA_NVCC_FLAGS="-DZKB_R1CS_MIN_CTAS=$m" python -c "import __graft_entry__ as g; g.build(force=True)" >/dev/null 2>&1
  echo "min_ctas=$m single: $(python scripts/r1cs_once.py 22 1 | tail -1)"
  echo "min_ctas=$m batch64: $(python scripts/r1cs_once.py 18 64 | tail -1)"
done
