#!/bin/bash
# C2 (single-witness Goldilocks, one launch for all wavefronts): counter barrier vs cooperative_groups grid.sync, CTA count
for v in "" "ZKB_CG_GRID_SYNC=1" "ZKB_COOP_BLOCKS=148" "ZKB_COOP_BLOCKS=296" "ZKB_CG_GRID_SYNC=1 ZKB_COOP_BLOCKS=148"; do
  echo "== $v"
  env $v timeout 300 python tests/bench_configs.py --only c2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['config'], '| ms', round(d['ms'], 4), 'us/level', round(d['us_per_level'], 3), 'barrier', {k: round(v, 3) for k, v in d['barrier'].items()}, 'floor frac', round(d['frac_of_barrier_floor'], 3))"
done
