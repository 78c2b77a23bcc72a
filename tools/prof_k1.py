"""ncu target (not a benchmark): K1 at (n, D) in a chosen form, `reps` launches.
    python tools/prof_k1.py 20 50000000 3 [reps]     # 3 = centred-Gram kernel, 2 = direct TMA-staged kernel"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops  # noqa: E402

n, D, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
X = torch.empty(n, D, device=dev)
for i in range(n):
    X[i].normal_(0.0, 0.05 * (1 + 0.1 * i), generator=g)
sc = ops.SvgdScratch.allocate(n, dev)
_lib.get().bde_tune(b"pairdist_variant", variant)
for _ in range(reps):
    ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)
torch.cuda.synchronize()
