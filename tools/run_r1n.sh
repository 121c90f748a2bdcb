mkdir -p gpurun_out
P=29617
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r1n_bench_c2_n4.json 2> gpurun_out/r1n_bench_c2_n4.err; cut -c1-260 gpurun_out/r1n_bench_c2_n4.json; wc -l gpurun_out/r1n_bench_c2_n4.json; tail -2 gpurun_out/r1n_bench_c2_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 4 --impl reference --steps 3 --warmup 1 > gpurun_out/r1n_ref_n4.json 2> gpurun_out/r1n_ref_n4.err; cut -c1-200 gpurun_out/r1n_ref_n4.json; wc -l gpurun_out/r1n_ref_n4.json
