#!/bin/bash
# r02 session 20 (8 GPUs): bench.py at N = 8 with the K1 -> K2 chain (weak line + strong scaling at D_total = 1e8 / 1e9)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
grep "\[bench\]" gpurun_out/r02_bench_n8.err | tail -12; head -c 600 gpurun_out/r02_bench_n8.json
