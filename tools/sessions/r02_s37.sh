#!/bin/bash
# r02 session 37: bench at N = 1 with the whole-step section (C1 / C2 with real model closures inside the driver-run line)
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/r02_bench_n1_s37.json 2> gpurun_out/r02_bench_n1_s37.err; echo "bench rc=$?"; grep "whole" gpurun_out/r02_bench_n1_s37.err | tail -4 | cut -c1-900; head -c 250 gpurun_out/r02_bench_n1_s37.json
