#!/bin/bash
# last validation of HEAD: GPU parity suite, smoke, whole-step harness (host-overhead changes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 400 python tests/perf_whole_step.py > gpurun_out/whole_step.json 2> gpurun_out/whole_step.err; echo "whole rc=$?"; tail -n 7 gpurun_out/whole_step.err | cut -c1-260
