#!/bin/bash
# r02 session 7 (2 GPUs): sharding / peer-exchange tests incl. the abandoned-exchange and rank-local cases; bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_gloo.py -q -m gpu > gpurun_out/r02_pytest_gpu_n2.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"
grep -E "strong|e2e skipped|Error|error" gpurun_out/r02_bench_n2.err | tail -8
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['exchange']); print(d['strong']); print(d['e2e'])"
