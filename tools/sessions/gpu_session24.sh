#!/bin/bash
# staged K2 with tile sets at n = 16 / 20: parity tests, then direct vs TS=3 vs TS=4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "tile_sets or variants or fused_apply" > gpurun_out/pytest_ts.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ts.log
tail -n 15 gpurun_out/pytest_ts.log
timeout 600 python tools/exp_tilesets.py > gpurun_out/exp_tilesets.jsonl 2> gpurun_out/exp_tilesets.err; echo "exp rc=$?"
cat gpurun_out/exp_tilesets.jsonl; tail -n 5 gpurun_out/exp_tilesets.err
