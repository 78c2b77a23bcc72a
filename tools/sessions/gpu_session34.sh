#!/bin/bash
# validation of HEAD: full parity suite, bench + reference arm, whole-step harness incl. C5 (eager reference order on the GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python tests/perf_whole_step.py > gpurun_out/whole_step.json 2> gpurun_out/whole_step.err; echo "whole rc=$?"; cat gpurun_out/whole_step.err | tail -n 8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
