set -u
out=gpurun_out; tag=r02a; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/${tag}_smi.log; free -g >> $out/${tag}_smi.log; nproc >> $out/${tag}_smi.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_gpu_tests.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 300 python bench.py --witnesses 256 --steps 1 --warmup 1 --no-cpu-baseline --no-value-check > $out/${tag}_plain256.json 2>/dev/null && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --witnesses 256 --steps 1 --warmup 1 --no-cpu-baseline --no-value-check > $out/${tag}_ncu_list.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_level_pipe -s 20 -c 3 -o $out/${tag}_prof_level \
      python bench.py --witnesses 256 --steps 1 --warmup 1 --no-cpu-baseline --no-value-check > $out/${tag}_ncu_full.log 2>&1
}
tail -3 $out/${tag}_gpu_tests.log; cat $out/${tag}_bench_n1.json; tail -5 $out/${tag}_bench_n1.err
