#!/bin/bash
# batched SWAG sampling: parity + timing; ncu of the small-D (C2 / C1) SVGD kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
grep "swag\|failed" gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'svgd' -c 18 -f -o gpurun_out/prof_small python tools/prof_small.py > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
