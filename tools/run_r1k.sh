mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r1k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1k_pytest.log
grep -n "^E  \|FAILED\|passed\|failed\|rc=" gpurun_out/r1k_pytest.log | head -30
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/r1k_bench_c1.json 2> gpurun_out/r1k_bench_c1.err; cut -c1-330 gpurun_out/r1k_bench_c1.json; tail -3 gpurun_out/r1k_bench_c1.err
