#!/bin/bash
# Multi-GPU measurements of the three sharding schemes on ONE box with G GPUs:
#   C3  1,024 chains split G ways (no collective)          -- bench.py default workload
#   C4  4,194,304 rows sharded G ways, L = 50, one NCCL all-reduce per gradient evaluation (its time broken out)
#   C5  20,480 stored samples split G ways, 1,048,576 test rows, one moment merge at the end
# usage: tools/run_multi_gpu.sh G [tag]     (results: gpurun_out/<tag>_*_n<G>.json)
G=${1:-2}; TAG=${2:-r2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541"
if [ "$G" = "1" ]; then RUN="python"; fi
$RUN bench.py --gpus $G --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_n$G.json 2> gpurun_out/${TAG}_bench_c3_n$G.err
$RUN bench.py --gpus $G --workload c4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4_n$G.json 2> gpurun_out/${TAG}_bench_c4_n$G.err
$RUN bench.py --gpus $G --workload c5 --steps 2 --warmup 1 --pred-samples $((20480 / G / 4)) > gpurun_out/${TAG}_bench_c5_n$G.json 2> gpurun_out/${TAG}_bench_c5_n$G.err
for w in c3 c4 c5; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${w}_n$G.json"))
    r = d.get("roofline", {})
    print("$w G=$G value %.4g %s e2e %.4g ms/step %.1f frac %.3f allreduce %s" % (d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], r.get("frac", 0), r.get("allreduce")))
except Exception as e:
    print("$w G=$G failed:", e); print(open("gpurun_out/${TAG}_bench_${w}_n$G.err").read()[-1500:])
PY
done
