#!/bin/bash
# A/B of the R1CS check: lazy reduction (one Montgomery reduction per linear combination) on / off x CTAs per SM (register
# cap 64 / 85).  Run on the GPU box; parity first (field primitives + R1CS tests with the default build).
python -m pytest tests/test_gpu_field.py tests/test_gpu_r1cs.py -x -q -m gpu 2>&1 | tail -3
for cfg in "1 4" "1 3" "0 4" "0 3"; do
  set -- $cfg
  ZKB_EXTRA_NVCC_FLAGS="-DZKB_R1CS_LAZY=$1 -DZKB_R1CS_MIN_CTAS=$2" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  echo "LAZY=$1 MIN_CTAS=$2"
  ZKB_DEBUG=1 python scripts/r1cs_once.py 22 1 2>&1 | grep -v "^$" | tail -2
  python scripts/r1cs_once.py 18 64 2>&1 | tail -1
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
