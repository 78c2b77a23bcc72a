#!/bin/bash
# r02 session 16: two-phase split-K (reduce kernel behind the product kernel) for BBBLinear / Rank1Linear: parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bbb_linear.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/exp_bbb_linear.py > gpurun_out/r02_bbb_linear.jsonl 2> gpurun_out/r02_bbb_linear.err; tail -3 gpurun_out/r02_bbb_linear.err; cat gpurun_out/r02_bbb_linear.jsonl
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_bbb_linear.py -m gpu -x -q -k "16-768-768 or 130-64-10 or 256-1024" 2>&1 | tail -5
