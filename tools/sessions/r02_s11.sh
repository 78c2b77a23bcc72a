#!/bin/bash
# r02 session 11: ncu evidence for the round-2 kernels (never a bench value): launch list of the bench step, full captures of
# K1 / K2 at n = 10 and K2 at n = 20, hardware-counter sections of the Gram kernel (its setmaxnreg hand-over does not survive
# ncu's SASS patching: a --set full capture hung in session 3), full capture of the tcgen05 BBBLinear kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'svgd_(apply|pairdist)_tma' -s 6 -c 2 -f -o gpurun_out/r02_prof_n10 python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_n10.log 2>&1; echo "ncu n10 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'svgd_apply_tma' -s 2 -c 2 -f -o gpurun_out/r02_prof_n20_k2 python tools/prof_svgd.py 20 50000000 2 > gpurun_out/r02_ncu_n20.log 2>&1; echo "ncu n20 K2 rc=$?"
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis"
timeout 300 ncu $SECS --clock-control none -k regex:'pairgram' -s 1 -c 1 -f -o gpurun_out/r02_prof_gram20 python tools/prof_k1.py 20 50000000 3 > gpurun_out/r02_ncu_gram20.log 2>&1; echo "ncu gram20 rc=$?"
timeout 300 ncu $SECS --clock-control none -k regex:'pairgram' -s 1 -c 1 -f -o gpurun_out/r02_prof_gram16 python tools/prof_k1.py 16 60000000 3 > gpurun_out/r02_ncu_gram16.log 2>&1; echo "ncu gram16 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bbb_linear' -s 2 -c 1 -f -o gpurun_out/r02_prof_bbb_linear python tools/prof_bbb_linear.py > gpurun_out/r02_ncu_bbl.log 2>&1; echo "ncu bbb_linear rc=$?"
ls -la gpurun_out/*.ncu-rep
