#!/bin/bash
# r02 session 25 (2 GPUs): D-sharded SWAG / iVON / BBB classes over NCCL, the SVGD sharding tests, bench at N = 2 with the
# sharded elementwise section
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_posteriors.py tests/test_sharding_gloo.py -m gpu -x -q > gpurun_out/r02_pytest_gpu_n2.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_n2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"; grep "sharded elementwise\|strong" gpurun_out/r02_bench_n2.err | tail -4 | cut -c1-1500; head -c 400 gpurun_out/r02_bench_n2.json
