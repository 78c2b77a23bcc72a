#!/bin/bash
# r02 session 31 (8 GPUs): bench.py at N = 8 on the final tree (weak line, strong scaling at D_total = 1e8 / 1e9 after the K1-tail
# work, sharded elementwise family)
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"
grep "\[bench\] strong\|sharded elementwise" gpurun_out/r02_bench_n8.err | tail -3 | cut -c1-1800; head -c 500 gpurun_out/r02_bench_n8.json
