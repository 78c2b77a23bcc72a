#!/bin/bash
# r02 session 39: smoke() and a slice of the kernel parity tests on the relinked library of the last tree
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke_final.txt
timeout 60 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "svgd_kernels_vs_oracle or swag_update_and_sample or ivon" 2>&1 | tail -3 | tee -a gpurun_out/r02_smoke_final.txt
