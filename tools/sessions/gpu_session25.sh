#!/bin/bash
# Re-entry validation of HEAD on a fresh B200: full GPU parity suite, N=1 bench + reference arm, launch list,
# n = 16 / 20 kernel timings, ncu captures of the training-step kernel (n=10) and of K1 / K2 at n=20.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "launches rc=$?"
timeout 600 python tools/exp_tilesets.py > gpurun_out/exp_tilesets.jsonl 2> gpurun_out/exp_tilesets.err; echo "exp rc=$?"
cat gpurun_out/exp_tilesets.jsonl
timeout 300 python tools/sweep.py --n 20 --d 50000000 > gpurun_out/sweep_n20.txt 2> gpurun_out/sweep_n20.err; echo "sweep20 rc=$?"
head -n 16 gpurun_out/sweep_n20.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'svgd' -s 4 -c 4 -o gpurun_out/prof_n10_all python tools/prof_svgd.py 10 100000000 > gpurun_out/ncu_n10.log 2>&1; echo "ncu10 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'svgd' -s 3 -c 3 -o gpurun_out/prof_n20_all python tools/prof_svgd.py 20 50000000 > gpurun_out/ncu_n20.log 2>&1; echo "ncu20 rc=$?"
ls -la gpurun_out
