#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'svgd' -f -o gpurun_out/prof_small python tools/prof_small.py > gpurun_out/ncu_small.out 2>&1
tail -n 3 gpurun_out/ncu_small.out
