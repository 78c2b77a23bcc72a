#!/bin/bash
# r02 session 32: K2 / K2f at small D, direct-LDG against the TMA-staged kernel (threshold of the auto rule)
mkdir -p gpurun_out
timeout 400 python tools/exp_apply_small.py > gpurun_out/r02_apply_small.jsonl 2> gpurun_out/r02_apply_small.err; echo "rc=$?"; cat gpurun_out/r02_apply_small.jsonl; tail -3 gpurun_out/r02_apply_small.err
