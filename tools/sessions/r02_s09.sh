#!/bin/bash
# r02 session 9: racecheck with every hazard printed (classified afterwards by tools/racecheck_classes.py), and racecheck of
# the direct-LDG variants (no async-proxy writes: everything racecheck can model)
mkdir -p gpurun_out
SEL='pairgram_vs_oracle or pairgram_guard or staged_apply_tile_sets or train_step_tile_geometries or staged_plain_apply'
timeout 900 compute-sanitizer --tool racecheck --print-limit 100000 --log-file gpurun_out/r02_racecheck_all.txt \
    python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$SEL" > gpurun_out/r02_racecheck_all_pytest.txt 2>&1
echo "racecheck(all hazards) rc=$?"; tail -2 gpurun_out/r02_racecheck_all_pytest.txt
python tools/racecheck_classes.py gpurun_out/r02_racecheck_all.txt > gpurun_out/r02_sanitizer_racecheck_classes.txt; cat gpurun_out/r02_sanitizer_racecheck_classes.txt
# direct-LDG kernels only (variant 1 forced by the tests' own parametrisation)
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_racecheck_direct.txt \
    python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "(test_pairdist_variants_agree_with_oracle and 1-) or (test_apply_variants_agree_with_oracle and 1-) or (fused_apply_base_optimizer_vs_oracle and 1-) or (train_step_next_distances and 1-) or swag_update_and_sample or prior_terms or kl_and_l2" > gpurun_out/r02_sanitizer_racecheck_direct_pytest.txt 2>&1
echo "racecheck(direct kernels) rc=$?"; tail -2 gpurun_out/r02_sanitizer_racecheck_direct_pytest.txt; tail -3 gpurun_out/r02_sanitizer_racecheck_direct.txt
rm -f gpurun_out/r02_racecheck_all.txt
