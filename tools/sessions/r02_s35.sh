#!/bin/bash
# r02 session 35 (2 GPUs): ColumnShardedModel closures over NCCL — D-sharded SWAG / iVON / SVGD on a real model against the plain classes
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_sharded_closure.py -m gpu -q -x -rs 2>&1 | tail -8 | tee gpurun_out/r02_pytest_sharded_closure_n2.txt
