#!/bin/bash
# r02 session 4: K2 family on tensor-map TMA loads: full GPU parity suite, then same-box A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_s04_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_s04_pytest.txt
for which in prev new; do
  if [ $which = prev ]; then export BDE_B200_LIB=$PWD/beyond_deep_ensembles_b200/lib/prev/libbde_b200.so; else unset BDE_B200_LIB; fi
  timeout 300 python tools/sweep_D.py --n 10 --D 100000000,300000000 --iters 10 > gpurun_out/r02_k2ab_${which}_n10.jsonl 2>> gpurun_out/r02_k2ab.err
  timeout 300 python tools/sweep_D.py --n 16 --D 60000000 --iters 10 > gpurun_out/r02_k2ab_${which}_n16.jsonl 2>> gpurun_out/r02_k2ab.err
  timeout 300 python tools/sweep_D.py --n 20 --D 50000000 --iters 10 > gpurun_out/r02_k2ab_${which}_n20.jsonl 2>> gpurun_out/r02_k2ab.err
  timeout 300 python tools/sweep_D.py --n 5 --D 200000000 --iters 10 > gpurun_out/r02_k2ab_${which}_n5.jsonl 2>> gpurun_out/r02_k2ab.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_k2ab_*.jsonl')):
    for l in open(f):
        r=json.loads(l)
        print(f.split('/')[-1][9:-6], {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('n','D','k1_ms','k2_ms','k2_frac','step_ms','step_frac','train_step_ms','train_step_frac')})
PY
tail -3 gpurun_out/r02_k2ab.err
