for b in default 148 74 32 296 592; do
  if [ $b = default ]; then unset ZKB_COOP_BLOCKS; else export ZKB_COOP_BLOCKS=$b; fi
  python tests/bench_configs.py --only c2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('blocks=$b', d['config'][:60], round(d['gates_per_s']/1e9,3),'G/s', round(d['ms'],4),'ms', round(d['us_per_level'],2),'us/level')"
done
