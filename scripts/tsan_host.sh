#!/bin/bash
# ThreadSanitizer over the threaded host passes (parse pool, bulk flat ingest, levelizer): a TSan build of libzkb.so under
# zkinterface-ir_b200/build_tsan/ and the host tests that drive those passes with several threads.  Run here (no GPU):
#   scripts/tsan_host.sh > profiles/rNN_tsan_host.log 2>&1
set -u
cd "$(dirname "$0")/.."
PKG=zkinterface-ir_b200
OUT=$PKG/build_tsan
mkdir -p $OUT
SAN="-fsanitize=thread,-fno-omit-frame-pointer,-g,-fPIC,-O1"
pids=()
for src in $PKG/csrc/*.cu $PKG/csrc/*.cpp; do
  obj=$OUT/$(basename $src).o
  if [ ! -f $obj ] || [ $src -nt $obj ] || [ -n "$(find $PKG/csrc include -newer $obj \( -name '*.h' -o -name '*.cuh' \) | head -1)" ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -Xcompiler $SAN -I include -c -o $obj $src &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fsanitize=thread -ldl -o $OUT/libzkb.so $OUT/*.o || exit 1
export ZKB_LIB_PATH=$PWD/$OUT/libzkb.so
export LD_PRELOAD="$(gcc -print-file-name=libtsan.so)"
export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0 suppressions=$PWD/scripts/tsan.supp"
export ZKB_PLAN_THREADS=6 ZKB_PARSE_THREADS=6
python -m pytest tests/test_flat_ingest_host.py tests/test_host_evaluator.py tests/test_call_groups.py tests/test_robustness_host.py tests/test_bulk_push_host.py \
  -x -q -m "not gpu" -p no:cacheprovider 2>&1 | tee /tmp/tsan_pytest.log | grep -E "WARNING: ThreadSanitizer|SUMMARY: ThreadSanitizer|passed|failed" | sort | uniq -c | sort -rn | head -20
python - <<'PY' 2>&1 | grep -E "WARNING: ThreadSanitizer|SUMMARY: ThreadSanitizer|plan ok" | sort | uniq -c | sort -rn | head -20
import sys, importlib
sys.path.insert(0, ".")
import zkb_loader
z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
for p, window in ((c.BLS12_381_FR, 0), (c.GOLDILOCKS, 4096)):
    circ = c.random_circuit(1 << 19, 256, p, 50, window=window)
    b = z.GpuBackend(-1)
    b.set_field(p)
    b.push_gates(circ.gates, circ.const_pool)
    b.finalize(False)
    print("plan ok", b.stats()["n_levels"])
PY
