#!/bin/bash
# r02 session 14: interleaved A/B of the ring depth (A = whole ring, B = capped) on the same build
mkdir -p gpurun_out
P=beyond_deep_ensembles_b200/lib/prev2/libbde_b200.so
for r in 128 160 96; do
  timeout 400 python tools/ab_libs.py --a $P --tune-b ring_kb=$r --shapes 10x100000000,16x60000000,5x200000000,8x100000000 --rounds 5 > gpurun_out/r02_ring_ab_$r.jsonl 2>> gpurun_out/r02_ring_ab.err
  echo "ring $r"; python - <<PY
import json
for l in open('gpurun_out/r02_ring_ab_$r.jsonl'):
    d=json.loads(l); print(d['n'], {k:(v['A_ms'],v['B_ms'],v['B_over_A']) for k,v in d.items() if isinstance(v,dict)})
PY
done
tail -2 gpurun_out/r02_ring_ab.err
