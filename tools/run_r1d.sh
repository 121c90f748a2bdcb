set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -k "wide or umma" -x -q > gpurun_out/r1d_wide.log 2>&1; echo "rc=$?" >> gpurun_out/r1d_wide.log
tail -30 gpurun_out/r1d_wide.log
timeout 120 python tools/sweep_prof.py 9600 > gpurun_out/r1d_sweep_c2.log 2>&1; cat gpurun_out/r1d_sweep_c2.log
timeout 120 python tools/sweep_prof.py 262144 > gpurun_out/r1d_sweep_c2l.log 2>&1; cat gpurun_out/r1d_sweep_c2l.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest.log
tail -5 gpurun_out/r1d_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1d_bench_c2.json 2> gpurun_out/r1d_bench_c2.err; cat gpurun_out/r1d_bench_c2.json; tail -3 gpurun_out/r1d_bench_c2.err
