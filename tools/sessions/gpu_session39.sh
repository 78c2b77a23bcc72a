#!/bin/bash
# ncu evidence for the elementwise family (incl. the batched samplers) and for the n = 20 kernels after the rework
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ivon|swag|l2_' -s 9 -c 9 -f -o gpurun_out/prof_ew python tools/prof_ew.py 2 > gpurun_out/ncu_ew.log 2>&1; echo "ncu ew rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'svgd' -s 3 -c 3 -f -o gpurun_out/prof_n20_all python tools/prof_svgd.py 20 50000000 > gpurun_out/ncu_n20.log 2>&1; echo "ncu20 rc=$?"
ls -la gpurun_out/*.ncu-rep
