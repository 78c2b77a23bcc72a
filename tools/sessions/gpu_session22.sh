#!/bin/bash
# re-entry check: GPU parity tests, default bench, then ncu of the n=20 kernels and the training-step kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 500 $NCU -k regex:svgd -s 4 -c 3 -f -o gpurun_out/prof_n20 python tools/prof_svgd.py 20 50000000 2 > gpurun_out/prof_n20.log 2>&1; echo "n20 rc=$?"
timeout 500 $NCU -k regex:svgd -s 4 -c 4 -f -o gpurun_out/prof_n10 python tools/prof_svgd.py 10 100000000 2 > gpurun_out/prof_n10.log 2>&1; echo "n10 rc=$?"
ls -la gpurun_out/*.ncu-rep
