#!/bin/bash
# r02 session 21: D-sharded SWAG / iVON / BBB classes + the elementwise / sampler parity tests after the Philox / Box-Muller
# instruction diet; batched samplers fast vs general; ncu --set full of the two fast batched samplers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_posteriors.py tests/test_gpu_kernels.py tests/test_optimizers.py -m gpu -x -q -k "sharded or shard or ivon or swag or gauss or philox or bbb or rank1 or kl or ensemble" > gpurun_out/r02_s21_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_s21_pytest.txt
timeout 300 python tools/exp_batch_samplers.py > gpurun_out/r02_batch_samplers.jsonl 2> gpurun_out/r02_batch_samplers.err; echo "exp rc=$?"; cat gpurun_out/r02_batch_samplers.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sample_batch_fast' -c 2 -f -o gpurun_out/r02_prof_batch python tools/exp_batch_samplers.py prof > gpurun_out/r02_prof_batch.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_prof_batch.log
