#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python tools/exp_small.py > gpurun_out/exp_small.txt 2>&1
grep -E "whole step|cap=0" gpurun_out/exp_small.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 25 gpurun_out/bench.err
