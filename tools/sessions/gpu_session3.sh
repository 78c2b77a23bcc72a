#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py > gpurun_out/sweep.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/ncu_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgd_pairdist -s 3 -c 1 -o gpurun_out/prof_pairdist_tma python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_pairdist.out 2>&1
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/bench.err
