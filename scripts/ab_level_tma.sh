#!/bin/bash
# A/B: k_level_pipe (per-thread 16-byte loads) vs k_level_tma (cp.async.bulk operand rows into a shared-memory ring) on the
# headline shape, 1024 witnesses (4 tiles of 256).  Values are checked against the oracle in every run (values_checked).
run() { python bench.py --witnesses 1024 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value']/1e9,'G gate-evals/s  frac',d['roofline']['frac'],'avg_launch_ms',d['roofline']['avg_launch_ms'],'values_checked',d.get('values_checked'),'clk',d['clocks']['sm_mhz'])"; }
echo "pipe"; run
for st in 3 4; do
  for per in 16 64; do
    ctas=$(( st == 3 ? 4 : 3 ))
    ZKB_EXTRA_NVCC_FLAGS="-DZKB_TMA_STAGES=$st -DZKB_TMA_MIN_CTAS=$ctas" python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
    echo "tma stages=$st ctas/SM=$ctas grid/SM=$per"; ZKB_LEVEL_TMA=1 ZKB_TMA_GRID_PER_SM=$per run
  done
done
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
echo "pipe again"; run
