#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tests/perf_whole_step.py --steps 5 --warmup 2 > gpurun_out/whole_step.json 2> gpurun_out/whole_step.err; echo "rc=$?"
tail -n 30 gpurun_out/whole_step.err
