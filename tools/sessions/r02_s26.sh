#!/bin/bash
# r02 session 26: where K1's fixed cost goes at small D (instrumented build, globaltimer stamps)
mkdir -p gpurun_out
BDE_B200_LIB=beyond_deep_ensembles_b200/lib/timing/libbde_b200.so timeout 300 python tools/exp_tail_timing.py > gpurun_out/r02_tail_timing.jsonl 2> gpurun_out/r02_tail_timing.err; echo "rc=$?"; cat gpurun_out/r02_tail_timing.jsonl; tail -3 gpurun_out/r02_tail_timing.err
