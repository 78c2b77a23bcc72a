#!/bin/bash
# 2-GPU session: peer-exchange parity tests, then the N=2 bench (peer form vs NCCL form)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_sharding_gloo.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_peer.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_peer.log
tail -n 30 gpurun_out/pytest_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -n 5 gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d['exchange'], d['kernels'], d['e2e'])"
