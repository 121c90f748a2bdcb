mkdir -p gpurun_out
timeout 200 python tools/usweep_debug.py 2>&1 | tail -40
timeout 120 python tools/sweep_prof.py 9600 2>&1 | tail -2
timeout 120 python tools/sweep_prof.py 262144 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r1e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest.log
grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r1e_pytest.log | head -30
