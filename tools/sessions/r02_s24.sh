#!/bin/bash
# r02 session 24: rows in flight (KU) of the 16-draw SWAG pass; ncu --set full of the SVGD step at the small configs (C2 / C1)
mkdir -p gpurun_out
timeout 300 python tools/exp_batch_samplers.py > gpurun_out/r02_batch_samplers_ku.jsonl 2> gpurun_out/r02_batch_samplers_ku.err; echo "exp rc=$?"; grep "ku\|default\|x1" gpurun_out/r02_batch_samplers_ku.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'svgd' -c 12 -f -o gpurun_out/r02_prof_small python tools/prof_small.py > gpurun_out/r02_prof_small.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_prof_small.log
timeout 300 python tools/exp_small.py > gpurun_out/r02_small_breakdown.txt 2>&1; echo "small rc=$?"; head -20 gpurun_out/r02_small_breakdown.txt
