#!/bin/bash
# r02 session 40: the GradScaler paths of the drop-in classes on the CUDA library after the D-sharded guard went into step()
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_optimizers.py -m gpu -q -x -k "scaler or amp or nan or svgd_steps or swag_matches or ivon_matches or bbb_rank1" 2>&1 | tail -3 | tee gpurun_out/r02_pytest_scaler_paths.txt
