#!/bin/bash
# A/B: early-clobber accumulator operands in the PTX chains (field_ptx.cuh) on the headline kernel; 1024 witnesses, alternating
for rep in 1 2 3; do
  for fl in "" "-DZKB_NO_EARLY_CLOBBER"; do
    ZKB_EXTRA_NVCC_FLAGS="$fl" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
    echo -n "flags[$fl] "
    python bench.py --witnesses 1024 --steps 3 --warmup 2 --no-cpu-baseline --no-value-check 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'] / 1e9, 2), 'G gate-evals/s', d['roofline']['kernel'], d['clocks']['sm_mhz'])"
  done
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
