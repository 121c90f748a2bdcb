set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1f_pytest.log
tail -3 gpurun_out/r1f_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1f_bench_c2.json 2> gpurun_out/r1f_bench_c2.err
for w in c1 c3; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_$w.json 2> gpurun_out/r1f_bench_$w.err; done
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --leapfrog 4 --no-cpu-baseline > gpurun_out/r1f_bench_c4_L4.json 2> gpurun_out/r1f_bench_c4.err
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r1f_bench_c5.json 2> gpurun_out/r1f_bench_c5.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f_smoke.log 2>&1; tail -1 gpurun_out/r1f_smoke.log
