#!/bin/bash
# strong scaling at N=2, measured: D_total = 1e8 and 1e9 split over two GPUs (kernel-only lines, --skip-extras)
mkdir -p gpurun_out
for DPG in 50000000 500000000; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --d-per-gpu $DPG --skip-extras > gpurun_out/bench_n2_strong_$DPG.json 2> gpurun_out/bench_n2_strong_$DPG.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_strong_$DPG.json')); print($DPG, d['value'], d['ms_per_step'], d['kernels'])"
done
