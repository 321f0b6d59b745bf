#!/bin/bash
# A/B of the R1CS ring kernel: look-ahead GA (gathers in flight per thread) x CTAs per SM (register cap).  Run on the GPU box.
for cfg in "3 3" "3 4" "2 4"; do
  set -- $cfg
  ZKB_EXTRA_NVCC_FLAGS="-DZKB_R1CS_GA=$1 -DZKB_R1CS_MIN_CTAS=$2" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "GA=$1 MIN_CTAS=$2"
  ZKB_DEBUG=1 python scripts/r1cs_once.py 22 1 2>&1 | grep -v "^$" | tail -2
  python scripts/r1cs_once.py 18 64 2>&1 | tail -1
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
