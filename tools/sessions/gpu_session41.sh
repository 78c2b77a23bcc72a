#!/bin/bash
# PCIe link rates of the box vs the e2e host-buffer step at several chunk sizes
mkdir -p gpurun_out
timeout 400 python tools/exp_pcie.py > gpurun_out/exp_pcie.jsonl 2> gpurun_out/exp_pcie.err; echo "rc=$?"
cat gpurun_out/exp_pcie.jsonl; tail -n 3 gpurun_out/exp_pcie.err
