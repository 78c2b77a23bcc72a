#!/bin/bash
# 4-GPU session: the N=4 bench exactly as the driver launches it (peer-exchange parity at world 4 ran in the previous call)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n4.txt 2>&1
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s"
tail -n 5 gpurun_out/bench_n4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n4.json')); print(d['value'], d['ms_per_step'], d['exchange'], d['kernels'], d['e2e'])"
