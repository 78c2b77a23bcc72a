#!/bin/bash
# 2-GPU session: NCCL sharding parity + scaling bench at N=1,2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_sharding_gloo.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --skip-extras > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?" >> gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
tail -n 4 gpurun_out/pytest_multi.log gpurun_out/bench_n2.err
