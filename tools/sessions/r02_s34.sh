#!/bin/bash
# r02 session 34: last validation of the round — smoke, whole GPU suite, bench at N = 1 (after the 65,536-column switch of K2 / K2f)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_n1.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; grep "\[bench\]" gpurun_out/r02_bench_n1.err | grep "resnet20\|uci\|batch16\|ivon_sample:\|swag_update" | cut -c1-200; head -c 600 gpurun_out/r02_bench_n1.json
