set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1c_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1c_bench_c2.json 2> gpurun_out/r1c_bench_c2.err
for w in c1 c2l c3 c4 c5; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_$w.json 2> gpurun_out/r1c_bench_$w.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_c2.csv python bench.py --steps 2 --warmup 1 --leapfrog 100 --no-cpu-baseline > gpurun_out/r1c_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_wide2 -s 3 -c 1 -o gpurun_out/r1c_wide2_full python tools/sweep_prof.py > gpurun_out/r1c_ncu_full.log 2>&1
tail -3 gpurun_out/r1c_pytest.log; cat gpurun_out/r1c_bench_c2.json
