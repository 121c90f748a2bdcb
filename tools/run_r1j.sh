mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r1j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j_pytest.log
grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r1j_pytest.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_bench_c2.json 2> gpurun_out/r1j_bench_c2.err; cut -c1-330 gpurun_out/r1j_bench_c2.json; tail -3 gpurun_out/r1j_bench_c2.err
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_bench_c1.json 2> gpurun_out/r1j_bench_c1.err; cut -c1-330 gpurun_out/r1j_bench_c1.json; tail -3 gpurun_out/r1j_bench_c1.err
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --leapfrog 4 --no-cpu-baseline > gpurun_out/r1j_bench_c4.json 2> gpurun_out/r1j_bench_c4.err; cut -c1-330 gpurun_out/r1j_bench_c4.json; tail -3 gpurun_out/r1j_bench_c4.err
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/r1j_bench_c5.json 2> gpurun_out/r1j_bench_c5.err; cut -c1-330 gpurun_out/r1j_bench_c5.json; tail -3 gpurun_out/r1j_bench_c5.err
