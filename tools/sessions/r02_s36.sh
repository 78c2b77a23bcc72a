#!/bin/bash
# r02 session 36 (2 GPUs): bench at N = 2 with the sharded-closure section (all-gather weights / reduce-scatter gradients + SWAG on 1/N of the columns)
mkdir -p gpurun_out
timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_closure.json 2> gpurun_out/r02_bench_n2_closure.err; echo "bench rc=$?"; grep "sharded closure" gpurun_out/r02_bench_n2_closure.err | tail -2 | cut -c1-1800; head -c 300 gpurun_out/r02_bench_n2_closure.json
