#!/bin/bash
# r02 session 1: centred-Gram K1 — parity tests, A/B against the direct kernel, ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pairgram or pairdist_variants or svgd_kernels_vs_oracle" > gpurun_out/r02_s01_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_s01_pytest.txt
timeout 300 python tools/exp_k1.py --n 16,20 --D 50000000,100000000 > gpurun_out/r02_k1_ab.jsonl 2> gpurun_out/r02_k1_ab.err; echo "exp rc=$?"; cat gpurun_out/r02_k1_ab.jsonl; tail -5 gpurun_out/r02_k1_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pairgram' -s 2 -c 2 -f -o gpurun_out/r02_prof_gram python tools/prof_svgd.py 20 50000000 1 > gpurun_out/r02_ncu_gram.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
