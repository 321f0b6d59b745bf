#!/bin/bash
# A/B of the level kernel's resident CTAs per SM (register cap 85 -> 64): rebuilds libzkb.so on the box for each setting
for m in 3 4 3 4; do
  ZKB_EXTRA_NVCC_FLAGS="-DZKB_LEVEL_PIPE_MIN_CTAS=$m" python -c "import __graft_entry__ as g; g.build(force=True)" >/dev/null 2>&1
  python bench.py --witnesses 1024 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('min_ctas=$m', round(d['value']/1e9,2), 'G/s frac', round(d['roofline']['frac'],4), 'clk', d['clocks']['sm_mhz'])"
done
