#!/bin/bash
# r02 session 30 (4 GPUs): the multi-GPU tests up to world 4 — D-sharded SWAG / iVON / BBB classes over NCCL, SVGD sharding with
# the in-kernel peer exchange
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n4.txt 2>&1
timeout 1200 python -m pytest tests/test_sharded_posteriors.py tests/test_sharding_gloo.py -m gpu -q > gpurun_out/r02_pytest_gpu_n4.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu_n4.txt
