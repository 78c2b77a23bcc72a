#!/bin/bash
# r02 session 29: validation of the final tree — smoke, the whole GPU suite, both bench arms at N = 1, launch list, ncu --set full
# of K1 / K2 at n = 10 and of the fast batched samplers, whole-step harness
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_n1.txt
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 400 gpurun_out/r02_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; grep "\[bench\]" gpurun_out/r02_bench_n1.err | tail -45 | cut -c1-260; head -c 1200 gpurun_out/r02_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'svgd_(apply|pairdist)_tma' -s 6 -c 2 -f -o gpurun_out/r02_prof_n10 python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/r02_ncu_n10.log 2>&1; echo "ncu n10 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sample_batch_fast' -c 4 -f -o gpurun_out/r02_prof_batch python tools/exp_batch_samplers.py prof > gpurun_out/r02_prof_batch.log 2>&1; echo "ncu batch rc=$?"
timeout 1200 python tests/perf_whole_step.py > gpurun_out/r02_whole_step.json 2> gpurun_out/r02_whole_step.err; echo "whole rc=$?"; grep "whole_step\] C" gpurun_out/r02_whole_step.err | cut -c1-300
