#!/bin/bash
# fused unscale + inf-check gather (f2): kernel vs torch's own unscale kernel, classes under a CUDA GradScaler; full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
