#!/bin/bash
# packed Adam: parity suite; small-D breakdown; ncu of the n=10 kernels in their new geometry; whole-step harness
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python tools/exp_small.py > gpurun_out/exp_small.txt 2> gpurun_out/exp_small.err; echo "small rc=$?"
grep "n=20 D=273664\|n=10 D=512 " gpurun_out/exp_small.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac']); [print(k, round(v['ms'],4), round(v['GBps'],1)) for k,v in d['paths'].items() if 'ms' in v and 'svgd' in k]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'svgd' -s 4 -c 4 -f -o gpurun_out/prof_n10_all python tools/prof_svgd.py 10 100000000 > gpurun_out/ncu_n10.log 2>&1; echo "ncu10 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "launches rc=$?"
timeout 900 python tests/perf_whole_step.py > gpurun_out/whole_step.json 2> gpurun_out/whole_step.err; echo "whole rc=$?"; tail -n 3 gpurun_out/whole_step.err
