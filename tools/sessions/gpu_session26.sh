#!/bin/bash
# n=20 rolled j-loop (unroll 5) on 4 sets, n=16 unrolled on 3 sets; 3x256 vs 1x512 at n <= 12 for every form
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python tools/exp_tilesets.py > gpurun_out/exp_tilesets.jsonl 2> gpurun_out/exp_tilesets.err; echo "exp rc=$?"
cat gpurun_out/exp_tilesets.jsonl; tail -n 5 gpurun_out/exp_tilesets.err
timeout 300 python tools/sweep.py --n 20 --d 50000000 > gpurun_out/sweep_n20.txt 2> gpurun_out/sweep_n20.err; echo "sweep20 rc=$?"
head -n 16 gpurun_out/sweep_n20.txt
timeout 300 python tools/sweep.py --n 16 --d 60000000 > gpurun_out/sweep_n16.txt 2> gpurun_out/sweep_n16.err; echo "sweep16 rc=$?"
head -n 16 gpurun_out/sweep_n16.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 4 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step']); [print(k, round(v['ms'],4), round(v['GBps'],1)) for k,v in d['paths'].items() if 'ms' in v]"
