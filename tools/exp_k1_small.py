#!/usr/bin/env python
"""Where does the centred-Gram K1 (n = 16 / 20) start to pay?  K1 (+ fused K1b) at small D with the L2 flushed between
launches (as between two training steps), forms: direct-LDG (1), direct TMA-staged (2), Gram (3).  JSON lines."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import time_kernel  # noqa: E402
from beyond_deep_ensembles_b200 import _lib, ops  # noqa: E402

lib = _lib.get()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (20, 16):
    for D in (273_664, 524_288, 1_048_576, 2_097_152, 4_194_304, 8_388_608):
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn(n, D, device=dev, generator=g) * 0.05
        sc = ops.SvgdScratch.allocate(n, dev)
        rec = {"n": n, "D": D}
        for name, v in (("direct_ldg", 1), ("direct_tma", 2), ("gram", 3)):
            lib.bde_tune(b"pairdist_variant", v)
            rec[name + "_us"] = round(1e3 * time_kernel(lambda: ops.svgd_pairdist_bandwidth(X, sc, 3e-4, 1.0, 50000.0), 20, 3, flush), 2)
        lib.bde_tune(b"pairdist_variant", 0)
        print(json.dumps(rec), flush=True)
