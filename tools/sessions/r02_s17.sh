#!/bin/bash
# r02 session 17: validation of the final tree — smoke, the whole GPU suite (incl. the live reference comparison), both
# bench arms at N = 1, the ncu launch list of the bench step
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_n1.txt
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 600 gpurun_out/r02_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; grep "\[bench\]" gpurun_out/r02_bench_n1.err | tail -40; head -c 1500 gpurun_out/r02_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_launches.log 2>&1; tail -2 gpurun_out/r02_launches.log | head -c 400
