#!/bin/bash
# r02 session 38: the one-rank-group ColumnShardedModel test on the CUDA library (runs in every 1-GPU suite)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_sharded_closure.py -m gpu -q -x -rs 2>&1 | tail -8 | tee gpurun_out/r02_pytest_sharded_closure_n1.txt
