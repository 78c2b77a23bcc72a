"""Experiment: 1 x 512 vs 3 x 256 consumer geometry of the staged K2 / training-step kernel at n = 10 as D grows
(same box, back to back).  One JSON line per (D, form, geometry)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops  # noqa: E402


def timed(fn, iters=8, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


lib = _lib.get()
dev = torch.device("cuda", 0)
n = 10
for D in (50_000_000, 100_000_000, 200_000_000, 300_000_000, 600_000_000):
    X = torch.randn(n, D, device=dev) * 0.05
    G = torch.randn(n, D, device=dev) * 1e-3
    out = torch.empty_like(X)
    s0 = torch.zeros(D, device=dev)
    sc = ops.SvgdScratch.allocate(n, dev)
    ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)
    nk = ops.NextKernel(True, 0.01, 1.0, 50000.0)
    sgd = dict(buf_initialized=True, lr=1e-7, momentum=0.9, nesterov=True, weight_decay=3e-4)
    forms = {"k2": (lambda: ops.svgd_apply(X, G, out, sc), 12.0 * n * D),
             "train_sgd": (lambda: ops.svgd_apply_sgd(X, G, sc, s0, next_kernel=nk, **sgd), (12.0 * n + 8.0) * D)}
    for form, (fn, nbytes) in forms.items():
        for label, ts in (("1x512", 1), ("3x256", 3), ("1x512", 1), ("3x256", 3)):
            lib.bde_tune(b"apply_variant", 2)
            lib.bde_tune(b"apply_tile_sets", ts)
            ms = timed(fn)
            lib.bde_tune(b"apply_variant", 0)
            lib.bde_tune(b"apply_tile_sets", 0)
            print(json.dumps({"n": n, "D": D, "form": form, "geometry": label, "ms": round(ms, 4),
                              "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
    del X, G, out, s0
    torch.cuda.empty_cache()
