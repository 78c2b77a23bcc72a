#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:svgd -s 3 -c 3 -f -o gpurun_out/prof_n20 python tools/prof_svgd.py 20 100000000 2 > gpurun_out/prof_n20.log 2>&1; echo "n20 rc=$?"
timeout 600 $NCU -k regex:svgd -s 4 -c 4 -f -o gpurun_out/prof_n10_300m python tools/prof_svgd.py 10 300000000 2 > gpurun_out/prof_n10_300m.log 2>&1; echo "n10 rc=$?"
ls -la gpurun_out/*.ncu-rep
