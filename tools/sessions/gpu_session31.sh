#!/bin/bash
# batched SWAG sampler after the load / occupancy rework: parity + A/B of draws per pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "swag" > gpurun_out/pytest_swag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_swag.log
tail -n 4 gpurun_out/pytest_swag.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
grep "swag\|failed" gpurun_out/bench.err
