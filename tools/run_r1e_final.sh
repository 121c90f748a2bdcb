set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest.log
tail -3 gpurun_out/r1e_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1e_bench_c2.json 2> gpurun_out/r1e_bench_c2.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r1e_ref_c2.json 2> gpurun_out/r1e_ref_c2.err
for w in c1 c2l c3; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_$w.json 2> gpurun_out/r1e_bench_$w.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches_c2.csv python bench.py --steps 2 --warmup 3 --leapfrog 100 --no-cpu-baseline > gpurun_out/r1e_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_wide2 -s 3 -c 1 -o gpurun_out/r1e_wide2_full python tools/sweep_prof.py > gpurun_out/r1e_ncu_full.log 2>&1
cut -c1-200 gpurun_out/r1e_bench_c2.json
