#!/bin/bash
# r02 session 3: Gram K1 with the producer warpgroup + setmaxnreg: parity tests, A/B, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "pairgram" > gpurun_out/r02_s03_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_s03_pytest.txt
timeout 300 python tools/exp_k1.py --n 16,20 --D 50000000,100000000 --iters 20 > gpurun_out/r02_k1_ab3.jsonl 2> gpurun_out/r02_k1_ab3.err; echo "exp rc=$?"; cut -c1-180 gpurun_out/r02_k1_ab3.jsonl; tail -5 gpurun_out/r02_k1_ab3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pairgram' -s 1 -c 1 -f -o gpurun_out/r02_prof_gram20b python tools/prof_k1.py 20 50000000 3 > gpurun_out/r02_ncu_gram20b.log 2>&1; echo "ncu rc=$?"
