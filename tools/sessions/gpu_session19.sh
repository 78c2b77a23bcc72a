#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 3 gpurun_out/bench.err
timeout 900 python tools/sweep_D.py > gpurun_out/sweep_D.jsonl 2> gpurun_out/sweep_D.err; echo "sweep rc=$?"
cat gpurun_out/sweep_D.jsonl; tail -n 5 gpurun_out/sweep_D.err
