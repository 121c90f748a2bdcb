mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -k "wide or umma or c2s" -x -q 2>&1 | tail -5
TBNN_US_PROF=1 timeout 120 python tools/sweep_prof.py 9600 2>&1 | tail -10
TBNN_US_PROF=1 timeout 120 python tools/sweep_prof.py 4736 2>&1 | tail -10
timeout 120 python tools/sweep_prof.py 9600 2>&1 | tail -1
timeout 120 python tools/sweep_prof.py 262144 2>&1 | tail -1
