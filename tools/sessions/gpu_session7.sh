#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ivon|swag|l2' -s 6 -c 6 -f -o gpurun_out/prof_ew python tools/prof_ew.py 2 > gpurun_out/ncu_ew.out 2>&1
tail -n 5 gpurun_out/ncu_ew.out
