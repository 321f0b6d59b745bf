#!/bin/bash
# The measurement pass behind profiles/ (run on a B200 box from the repo root; outputs under gpurun_out/).
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'scripts/measure_all.sh r01x'
# Every ncu capture follows a plain run of the same command that exited 0; numbers printed under ncu are never
# bench values.  Multi-GPU lines: torchrun --nproc-per-node N bench.py --gpus N (gpurun --gpus N).
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/${tag}_gpu_tests.log
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
# launch list + full capture of the dominant kernel (one 256-witness tile)
python bench.py --witnesses 256 --steps 1 --warmup 1 --no-cpu-baseline --no-value-check > $out/${tag}_plain256.json 2>/dev/null && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --witnesses 256 --steps 1 --warmup 1 --no-cpu-baseline --no-value-check > $out/${tag}_ncu_list.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_level_tma -s 20 -c 3 -o $out/${tag}_prof_level \
      python bench.py --witnesses 256 --steps 1 --warmup 1 --no-cpu-baseline --no-value-check > $out/${tag}_ncu_full.log 2>&1
}
# secondary configs (parity checked in the same run) and the R1CS kernel capture
python tests/bench_configs.py > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err
python tests/bench_configs.py --only c3s >> $out/${tag}_configs.jsonl 2>> $out/${tag}_configs.err
python scripts/r1cs_once.py 22 1 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_r1cs_check -c 2 -o $out/${tag}_prof_r1cs \
      python scripts/r1cs_once.py 22 1 > $out/${tag}_ncu_r1cs.log 2>&1
python tests/bench_configs.py --only c5 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_bool_groups -c 3 -o $out/${tag}_prof_groups \
      python tests/bench_configs.py --only c5 > $out/${tag}_ncu_groups.log 2>&1
python scripts/field_throughput.py > $out/${tag}_field_throughput.jsonl 2>/dev/null
for f in m31 goldilocks p124; do
  python bench.py --field $f --no-cpu-baseline --no-value-check --steps 2 --warmup 2 2>/dev/null > $out/${tag}_bench_$f.json
done
# then, here: python profiles/summarize_ncu.py gpurun_out/${tag}_prof_level.ncu-rep profiles/${tag}_k_level_pipe_ncu_full_summary.csv
#             python profiles/make_traffic_json.py profiles/${tag}_k_level_pipe_ncu_full_summary.csv 20 256
