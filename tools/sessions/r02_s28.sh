#!/bin/bash
# r02 session 28: host-time profile (cProfile) of step() at C1 / C4a / C2
mkdir -p gpurun_out
timeout 900 python tests/perf_whole_step.py --configs C1,C4a,C2 --cprofile 45 > gpurun_out/r02_whole_step_prof.json 2> gpurun_out/r02_whole_step_prof.err; echo "rc=$?"; grep "whole_step\] C" gpurun_out/r02_whole_step_prof.err
