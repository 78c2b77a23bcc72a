#!/bin/bash
# blocked pair decomposition of K1 at n = 16 / 20: parity suite, sweeps, small-D timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python tools/sweep.py --n 20 --d 50000000 > gpurun_out/sweep_n20.txt 2> gpurun_out/sweep_n20.err; echo "sweep20 rc=$?"
grep "pairdist\|svgd_step" gpurun_out/sweep_n20.txt
timeout 300 python tools/sweep.py --n 16 --d 60000000 > gpurun_out/sweep_n16.txt 2> gpurun_out/sweep_n16.err; echo "sweep16 rc=$?"
grep "pairdist\|svgd_step" gpurun_out/sweep_n16.txt
timeout 300 python tools/exp_small.py > gpurun_out/exp_small.txt 2> gpurun_out/exp_small.err; echo "small rc=$?"
grep "n=20 D=273664 cold=True" gpurun_out/exp_small.txt
