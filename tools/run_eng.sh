timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3', d['value'], d['roofline']['launch_ms'], d['roofline']['fp32_tflops'])"
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --leapfrog 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4', d['value'], d['roofline']['launch_ms'], d['roofline']['fp32_tflops'])"
