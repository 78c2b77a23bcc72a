#!/bin/bash
# K2 on 3x256 at n >= 8: full parity suite + bench + n<=12 geometry table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 4 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels']); [print(k, round(v['ms'],4), round(v['GBps'],1)) for k,v in d['paths'].items() if 'ms' in v]"
timeout 600 python tools/exp_tilesets.py > gpurun_out/exp_tilesets.jsonl 2> gpurun_out/exp_tilesets.err; echo "exp rc=$?"
