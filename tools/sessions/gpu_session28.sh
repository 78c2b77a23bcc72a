#!/bin/bash
# 2-GPU session at HEAD: sharding parity tests (NCCL + in-kernel peer exchange), the N=2 bench as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharding_gloo.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n2.log
tail -n 6 gpurun_out/pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -n 5 gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d.get('exchange'), d['kernels'], d['e2e'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_n2_ref.json
