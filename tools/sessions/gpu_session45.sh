#!/bin/bash
# sanity of the re-split instantiation files: smoke + the SVGD kernel parity tests (all n)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 200 -p no:cacheprovider -k "svgd_kernels_vs_oracle or variants_agree or fused_apply_base" > gpurun_out/pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_split.log
tail -n 3 gpurun_out/pytest_split.log
