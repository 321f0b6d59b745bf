#!/bin/bash
run() { python bench.py --witnesses 1024 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value']/1e9,'G gate-evals/s  frac',d['roofline']['frac'],'avg_launch_ms',d['roofline']['avg_launch_ms'],'values_checked',d.get('values_checked'),'clk',d['clocks']['sm_mhz'])"; }
for per in 32 64 128 256 1024; do
  echo "tma stages=3 ctas/SM=4 grid/SM=$per"; ZKB_LEVEL_TMA=1 ZKB_TMA_GRID_PER_SM=$per run
done
ZKB_EXTRA_NVCC_FLAGS="-DZKB_TMA_STAGES=2 -DZKB_TMA_MIN_CTAS=5" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
for per in 64 128; do
  echo "tma stages=2 ctas/SM=5 grid/SM=$per"; ZKB_LEVEL_TMA=1 ZKB_TMA_GRID_PER_SM=$per run
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
