#!/bin/bash
# r02 session 6: interleaved same-process A/B, previous build (per-row bulk copies in K2) against tensor-map K2, two ring depths
mkdir -p gpurun_out
P=beyond_deep_ensembles_b200/lib/prev/libbde_b200.so
timeout 500 python tools/ab_libs.py --a $P > gpurun_out/r02_ab_full.jsonl 2> gpurun_out/r02_ab.err; echo full ring; cat gpurun_out/r02_ab_full.jsonl
timeout 500 python tools/ab_libs.py --a $P --tune-b ring_kb=128 > gpurun_out/r02_ab_128.jsonl 2>> gpurun_out/r02_ab.err; echo ring 128; cat gpurun_out/r02_ab_128.jsonl
tail -3 gpurun_out/r02_ab.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
