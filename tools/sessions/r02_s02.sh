#!/bin/bash
# r02 session 2: ncu captures of K1 at n = 20 (centred-Gram and direct) and n = 16 (Gram)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pairgram' -s 1 -c 1 -f -o gpurun_out/r02_prof_gram20 python tools/prof_k1.py 20 50000000 3 > gpurun_out/r02_ncu_gram20.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pairdist_tma' -s 1 -c 1 -f -o gpurun_out/r02_prof_direct20 python tools/prof_k1.py 20 50000000 2 > gpurun_out/r02_ncu_direct20.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pairgram' -s 1 -c 1 -f -o gpurun_out/r02_prof_gram16 python tools/prof_k1.py 16 50000000 3 > gpurun_out/r02_ncu_gram16.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
