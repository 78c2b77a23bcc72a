#!/bin/bash
# C5 sweep (BASELINE configs[4]) with the current kernels: n = 5 / 10 / 20, D = 10 M ... 1 B on one GPU
mkdir -p gpurun_out
timeout 900 python tools/sweep_D.py > gpurun_out/sweep_D.jsonl 2> gpurun_out/sweep_D.err; echo "sweep rc=$?"
cut -c1-400 gpurun_out/sweep_D.jsonl; tail -n 3 gpurun_out/sweep_D.err
