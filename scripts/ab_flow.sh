#!/bin/bash
# A/B of the dataflow launch (k_levels_flow) on C1 / C2: off (barrier kernels), poll back-off 0 / 20 / 100 ns, CTAs taking part
for cfg in "ZKB_FLOW=0" "ZKB_FLOW=1 ZKB_FLOW_SLEEP=0" "ZKB_FLOW=1 ZKB_FLOW_MIN=1" "ZKB_FLOW=1 ZKB_FLOW_SLEEP=50" "ZKB_FLOW=1 ZKB_FLOW_BLOCKS=148" "ZKB_FLOW=1 ZKB_FLOW_BLOCKS=296" "ZKB_FLOW=1 ZKB_FLOW_BLOCKS=592"; do
  echo "== $cfg"
  env $cfg timeout 200 python tests/bench_configs.py --only c1,c2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  ', d['config'][:60], '| ms', round(d.get('ms', d.get('device_ms')), 4), 'us/level', round(d.get('us_per_level', 0), 3), '| wavefronts only: ms', round(d.get('levels_ms', 0), 4), 'us/level', round(d.get('us_per_level_levels_only', 0), 3))
"
done
