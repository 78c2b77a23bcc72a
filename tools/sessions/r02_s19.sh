#!/bin/bash
# r02 session 19: fused forms at n <= 12 — A = default geometry (1 x 512, whole ring), B = 3 x 256 tile sets / capped rings
mkdir -p gpurun_out
P=beyond_deep_ensembles_b200/lib/prevA/libbde_b200.so
: > gpurun_out/r02_fused_geom_ab.jsonl
for t in apply_tile_sets=3 ring_kb=128 ring_kb=160; do
  echo "{\"tune_b\": \"$t\"}" >> gpurun_out/r02_fused_geom_ab.jsonl
  timeout 400 python tools/ab_libs.py --a $P --tune-b $t --shapes 10x100000000,12x80000000,8x100000000 --rounds 5 >> gpurun_out/r02_fused_geom_ab.jsonl 2>> gpurun_out/r02_fused_geom_ab.err
done
tail -2 gpurun_out/r02_fused_geom_ab.err
python - <<PY
import json
for l in open('gpurun_out/r02_fused_geom_ab.jsonl'):
    d=json.loads(l)
    if 'tune_b' in d: print(d); continue
    print(d['n'], {k:(v['A_ms'],v['B_ms'],v['B_over_A']) for k,v in d.items() if isinstance(v,dict)})
PY
