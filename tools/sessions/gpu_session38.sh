#!/bin/bash
# validation after the stride-dependent geometry switch: full parity suite, n=10 D-sweep, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep_D.py --n 10 > gpurun_out/sweep_D_n10.jsonl 2> gpurun_out/sweep_D_n10.err; echo "sweep rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/sweep_D_n10.jsonl'):
    d=json.loads(l); print(d['n'], d['D'], {k: round(v,4) for k,v in d.items() if k.endswith('_ms') or k.endswith('_frac')})
PY
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
