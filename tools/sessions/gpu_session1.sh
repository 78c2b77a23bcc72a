#!/bin/bash
# First GPU session: smoke, parity tests, bench, sweep, ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os, psutil; print('cpus', os.cpu_count(), 'ram_gb', psutil.virtual_memory().total >> 30)" >> gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/sweep.py > gpurun_out/sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/ncu_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgd_apply -s 3 -c 1 -o gpurun_out/prof_apply python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_apply.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:svgd_pairdist -s 3 -c 1 -o gpurun_out/prof_pairdist python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_pairdist.out 2>&1
ls -la gpurun_out
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench.err
