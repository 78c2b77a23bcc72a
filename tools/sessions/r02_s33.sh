#!/bin/bash
# r02 session 33: staged K2 / K2f from 64 K columns on — SVGD parity tests (kernels, optimizers, live reference), small-D step times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_optimizers.py tests/test_reference_on_gpu.py tests/test_sharding_gloo.py -m gpu -x -q -k "svgd or pairdist or bandwidth or rbf or train or gram or tile or kernel or reference or apply" > gpurun_out/r02_s33_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_s33_pytest.txt
timeout 300 python tools/exp_small.py 2>&1 | grep "whole step" | head -12
