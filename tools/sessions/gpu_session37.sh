#!/bin/bash
# same-box A/B of the K2 consumer geometry as D grows (the D-sweep hinted that 3 x 256 loses beyond D = 3e8)
mkdir -p gpurun_out
timeout 600 python tools/exp_geometry_D.py > gpurun_out/exp_geometry_D.jsonl 2> gpurun_out/exp_geometry_D.err; echo "rc=$?"
cat gpurun_out/exp_geometry_D.jsonl; tail -n 3 gpurun_out/exp_geometry_D.err
