#!/bin/bash
# Instrumented build of the library (-DBDE_TAIL_TIMING) into beyond_deep_ensembles_b200/lib/timing/ for tools/exp_tail_timing.py.
set -e
cd "$(dirname "$0")/.."
OUT=beyond_deep_ensembles_b200/lib/timing; mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fno-gnu-unique -Iinclude -Ibeyond_deep_ensembles_b200/csrc -DBDE_TAIL_TIMING"
ls beyond_deep_ensembles_b200/csrc/*.cu | xargs -P 8 -I{} sh -c "nvcc $FLAGS -c {} -o $OUT/obj/\$(basename {} .cu).o"
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libbde_b200.so $OUT/obj/*.o
rm -rf $OUT/obj; ls -la $OUT
