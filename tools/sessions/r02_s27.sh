#!/bin/bash
# r02 session 27: K1 tail after the value-major partials and the warp-local sort barriers: stamps, SVGD parity tests, small-D step times
mkdir -p gpurun_out
BDE_B200_LIB=beyond_deep_ensembles_b200/lib/timing/libbde_b200.so timeout 300 python tools/exp_tail_timing.py > gpurun_out/r02_tail_timing_after.jsonl 2> gpurun_out/r02_tail_timing.err; echo "rc=$?"; cat gpurun_out/r02_tail_timing_after.jsonl; tail -3 gpurun_out/r02_tail_timing.err
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_optimizers.py tests/test_sharding_gloo.py -m gpu -x -q -k "svgd or pairdist or bandwidth or rbf or train or gram or tile or kernel" > gpurun_out/r02_s27_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_s27_pytest.txt
timeout 300 python tools/exp_small.py 2>&1 | grep "whole step\|K1b alone" | head -12
