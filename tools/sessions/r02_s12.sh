#!/bin/bash
# r02 session 12: bench (both arms) at N = 1 with the round-2 lines; whole-step harness C1-C5; f4 timing
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; grep -E "bench\]" gpurun_out/r02_bench_n1.err | tail -40
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/exp_bbb_linear.py > gpurun_out/r02_bbb_linear.jsonl 2> gpurun_out/r02_bbb_linear.err; cat gpurun_out/r02_bbb_linear.jsonl
timeout 1200 python tests/perf_whole_step.py > gpurun_out/r02_whole_step.json 2> gpurun_out/r02_whole_step.err; echo "whole rc=$?"; grep whole_step gpurun_out/r02_whole_step.err | tail -8
