#!/bin/bash
# A/B: cudaLimitMaxL2FetchGranularity (bytes fetched from DRAM per L2 miss) on the random-gather ceiling and the R1CS check.
for g in 64 32 128; do
  echo "L2_FETCH=$g"
  ZKB_L2_FETCH_GRANULARITY=$g ZKB_DEBUG=1 python scripts/r1cs_once.py 22 1 2>&1 | grep -v "^$" | tail -2
  ZKB_L2_FETCH_GRANULARITY=$g python scripts/r1cs_once.py 18 64 2>&1 | tail -1
  ZKB_L2_FETCH_GRANULARITY=$g python - <<'PY'
import zkb_loader
z = zkb_loader.load()
b = z.GpuBackend(0)
for mb in (168, 1024):
    print("gather", mb, "MB table:", b.debug_gather_throughput(mb << 20) / 1e9, "GB/s")
PY
done
