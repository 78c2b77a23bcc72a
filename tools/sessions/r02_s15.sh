#!/bin/bash
# r02 session 15: fused Rank1Linear forward — kernel parity, the live reference cases that contain the layer, timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bbb_linear.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_reference_on_gpu.py -m gpu -q -k "rank1 or bbb" 2>&1 | tail -15
timeout 300 python tools/exp_bbb_linear.py > gpurun_out/r02_bbb_linear.jsonl 2> gpurun_out/r02_bbb_linear.err; tail -3 gpurun_out/r02_bbb_linear.err; cat gpurun_out/r02_bbb_linear.jsonl
