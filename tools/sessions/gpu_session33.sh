#!/bin/bash
# 8-GPU session: the N=8 and N=4 bench exactly as the driver launches them (+ peer-exchange parity at world 4 / 8)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
timeout 600 python -m pytest tests/test_sharding_gloo.py -m gpu -q -x --timeout 400 -p no:cacheprovider > gpurun_out/pytest_n8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n8.log
tail -n 4 gpurun_out/pytest_n8.log
for N in 8 4; do
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$? wall=$(( $(date +%s) - T0 ))s"
tail -n 3 gpurun_out/bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('exchange'), d['e2e']['value'])"
done
