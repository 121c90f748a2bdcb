mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r1i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1i_pytest.log
grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r1i_pytest.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1i_bench_c2.json 2> gpurun_out/r1i_bench_c2.err; cat gpurun_out/r1i_bench_c2.json | cut -c1-400
