#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the C++ host side (reader, Evaluator mirror, recorder, levelizer,
# validator, stats, writer, host parts of the .cu files): a sanitized copy of libzkb.so is built under
# zkinterface-ir_b200/build_asan/ and the CPU test-suite runs against it (ZKB_LIB_PATH).  compute-sanitizer is closed on the
# GPU pool, so this is the memory-safety evidence SURVEY.md section 5 asks for on the host.  Run here (no GPU needed):
#   scripts/asan_host.sh > profiles/rNN_asan_host.log 2>&1
set -u
cd "$(dirname "$0")/.."
PKG=zkinterface-ir_b200
OUT=$PKG/build_asan
mkdir -p $OUT
SAN="-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer,-fno-sanitize-recover=undefined,-g,-fPIC,-O1"
pids=()
for src in $PKG/csrc/*.cu $PKG/csrc/*.cpp; do
  obj=$OUT/$(basename $src).o
  if [ ! -f $obj ] || [ $src -nt $obj ] || [ -n "$(find $PKG/csrc include -newer $obj \( -name '*.h' -o -name '*.cuh' \) | head -1)" ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -Xcompiler $SAN -I include -c -o $obj $src &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fsanitize=address,-fsanitize=undefined -ldl -o $OUT/libzkb.so $OUT/*.o || exit 1
export ZKB_LIB_PATH=$PWD/$OUT/libzkb.so
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:protect_shadow_gap=0
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider \
  --deselect tests/test_cpp_api.py --deselect tests/test_cli_host.py 2>&1 | tail -15
