#!/bin/bash
# r02 session 22: L2 prefetch in the fast batched samplers (A/B by distance), iVON passes with a runtime draw count
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ivon or swag" > gpurun_out/r02_s22_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_s22_pytest.txt
timeout 300 python tools/exp_batch_samplers.py > gpurun_out/r02_batch_samplers.jsonl 2> gpurun_out/r02_batch_samplers.err; echo "exp rc=$?"; cat gpurun_out/r02_batch_samplers.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sample_batch_fast' -c 4 -f -o gpurun_out/r02_prof_batch python tools/exp_batch_samplers.py prof > gpurun_out/r02_prof_batch.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_prof_batch.log
