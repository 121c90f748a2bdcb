#!/bin/bash
# round-2 closing pass on one GPU (final kernels): GPU test suite, C3 / C4 / C5 bench lines, launch list, full ncu capture of
# the 64-wide training sweep, memcheck of the two training-sweep variants (results: gpurun_out/r2j_*)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -3 gpurun_out/r2j_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r2j_bench_c5.json 2> gpurun_out/r2j_bench_c5.err
for w in c3 c4 c5; do cut -c1-220 gpurun_out/r2j_bench_$w.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_c3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2j_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_train_umma -s 2 -c 1 -o gpurun_out/r2j_tu_c3 python tools/tu_time.py c3 0 > gpurun_out/r2j_ncu_c3.log 2>&1
for c in train_umma64 train_umma128; do
  timeout 120 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize.py $c > gpurun_out/r2j_sanitizer_memcheck_$c.log 2>&1
  echo "rc=$? tool=memcheck case=$c" >> gpurun_out/r2j_sanitizer_memcheck_$c.log; tail -2 gpurun_out/r2j_sanitizer_memcheck_$c.log | tr '\n' ' '; echo
done
