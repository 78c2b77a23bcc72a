#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_gloo.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench2 rc=$?" >> gpurun_out/bench_n2.log
tail -n 3 gpurun_out/pytest_multi.log; tail -c 3000 gpurun_out/bench_n2.log
