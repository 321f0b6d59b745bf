#!/bin/bash
# A/B: bulk stores through shared memory on top of the bulk loads
run() { python bench.py --witnesses 1024 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value']/1e9,'G gate-evals/s  frac',d['roofline']['frac'],'avg_launch_ms',d['roofline']['avg_launch_ms'],'values_checked',d.get('values_checked'),'clk',d['clocks']['sm_mhz'])"; }
echo "tma loads only, stages=3 ctas/SM=4 (default build)"; ZKB_LEVEL_TMA=1 run
for cfg in "2 4" "3 3"; do
  set -- $cfg
  ZKB_EXTRA_NVCC_FLAGS="-DZKB_TMA_STORE=1 -DZKB_TMA_STAGES=$1 -DZKB_TMA_MIN_CTAS=$2" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
  ZKB_LEVEL_TMA=1 python -m pytest tests/test_gpu_flat.py -m gpu -x -q 2>&1 | tail -1
  for per in 64 128; do
    echo "tma loads+stores stages=$1 ctas/SM=$2 grid/SM=$per"; ZKB_LEVEL_TMA=1 ZKB_TMA_GRID_PER_SM=$per run
  done
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
