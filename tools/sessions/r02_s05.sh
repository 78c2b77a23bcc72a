#!/bin/bash
# r02 session 5: ring-depth sweep of the tensor-map kernels against the previous (per-row bulk copy) build
mkdir -p gpurun_out
export BDE_B200_LIB=$PWD/beyond_deep_ensembles_b200/lib/prev/libbde_b200.so
timeout 400 python tools/exp_ring.py --rings 0 --iters 20 > gpurun_out/r02_ring_prev.jsonl 2> gpurun_out/r02_ring.err; echo prev; cat gpurun_out/r02_ring_prev.jsonl
unset BDE_B200_LIB
timeout 600 python tools/exp_ring.py --iters 20 --rings 0,160,128,96 > gpurun_out/r02_ring_new.jsonl 2>> gpurun_out/r02_ring.err; echo new; cat gpurun_out/r02_ring_new.jsonl
tail -3 gpurun_out/r02_ring.err
