#!/bin/bash
# r02 session 41: every class-level GPU test (tests/test_optimizers.py) on the last tree
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_optimizers.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r02_pytest_optimizers_final.txt
