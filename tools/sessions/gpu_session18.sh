#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log
timeout 1200 python tests/perf_whole_step.py --steps 5 --warmup 2 > gpurun_out/whole_step.json 2> gpurun_out/whole_step.err; echo "rc=$?"
tail -n 8 gpurun_out/whole_step.err
