#!/bin/bash
# r02 session 18: K2f at n = 16 — A = unrolled j-loop on 3 tile sets (previous build), B = rolled on 4 sets, B' = rolled on 3 sets;
# parity of the changed kernels
mkdir -p gpurun_out
P=beyond_deep_ensembles_b200/lib/prevA/libbde_b200.so
timeout 400 python tools/ab_libs.py --a $P --shapes 16x60000000,16x20000000 --rounds 5 > gpurun_out/r02_k2f_n16_ab.jsonl 2> gpurun_out/r02_k2f_n16_ab.err
timeout 300 python tools/ab_libs.py --a $P --tune-b apply_tile_sets=3 --shapes 16x60000000 --rounds 5 >> gpurun_out/r02_k2f_n16_ab.jsonl 2>> gpurun_out/r02_k2f_n16_ab.err
tail -2 gpurun_out/r02_k2f_n16_ab.err; cat gpurun_out/r02_k2f_n16_ab.jsonl
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "apply or fused or sgd or adam or tile" 2>&1 | tail -4
