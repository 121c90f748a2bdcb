#!/bin/bash
# round-2 final single-GPU pass: GPU test suite, bench lines of every workload, launch list + full ncu capture of the
# training sweep at C3 and C4 (results: gpurun_out/r2i_*)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench_c3.json 2> gpurun_out/r2i_bench_c3.err
cut -c1-400 gpurun_out/r2i_bench_c3.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_ref_c3.json 2> gpurun_out/r2i_ref_c3.err
timeout 400 python bench.py --workload c4 --steps 3 --warmup 1 > gpurun_out/r2i_bench_c4.json 2> gpurun_out/r2i_bench_c4.err
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/r2i_bench_c1.json 2> gpurun_out/r2i_bench_c1.err
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r2i_bench_c5.json 2> gpurun_out/r2i_bench_c5.err
for w in c4 c2 c1 c5; do cut -c1-260 gpurun_out/r2i_bench_$w.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches_c3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_train_umma -s 2 -c 1 -o gpurun_out/r2i_tu_c3 python tools/tu_time.py c3 0 > gpurun_out/r2i_ncu_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_train_umma -s 2 -c 1 -o gpurun_out/r2i_tu_c4 python tools/tu_time.py c4 0 > gpurun_out/r2i_ncu_c4.log 2>&1
ls -la gpurun_out/r2i_*
