#!/bin/bash
# r02 session 10: f2 A/B (pre-bound gradient views against the gather launch) with real closures at C1 / C2 / C4b; where the Gram K1 starts to pay
mkdir -p gpurun_out
timeout 300 python tools/exp_k1_small.py > gpurun_out/r02_k1_small.jsonl 2> gpurun_out/r02_k1_small.err; cat gpurun_out/r02_k1_small.jsonl; tail -2 gpurun_out/r02_k1_small.err
for pb in on off; do
  timeout 900 python tests/perf_whole_step.py --configs C1,C2,C4b --prebind $pb > gpurun_out/r02_whole_step_prebind_$pb.json 2> gpurun_out/r02_whole_step_prebind_$pb.err; echo "prebind $pb rc=$?"
  grep -E "whole_step" gpurun_out/r02_whole_step_prebind_$pb.err | tail -8
done
