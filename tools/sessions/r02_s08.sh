#!/bin/bash
# r02 session 8: compute-sanitizer (memcheck / racecheck / synccheck) over the staged kernels: mbarrier rings with multi-set
# stage hand-over, ticket reduction, tensor-map TMA, setmaxnreg register shift, training-step tail
mkdir -p gpurun_out
SEL='pairgram_vs_oracle or pairgram_guard or staged_apply_tile_sets or train_step_tile_geometries or test_pairdist_variants_agree_with_oracle or staged_plain_apply'
for tool in memcheck synccheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_$tool.txt \
      python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$SEL" > gpurun_out/r02_sanitizer_${tool}_pytest.txt 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r02_sanitizer_${tool}_pytest.txt; tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
