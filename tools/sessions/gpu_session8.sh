#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "swag or ivon" > gpurun_out/pytest_ew.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ew.log
timeout 600 python tools/exp_ew.py > gpurun_out/exp_ew.txt 2>&1
tail -n 15 gpurun_out/pytest_ew.log; cat gpurun_out/exp_ew.txt
