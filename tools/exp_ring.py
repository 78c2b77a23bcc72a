#!/usr/bin/env python
"""Ring-depth sweep of the staged SVGD kernels (tuning knob ring_kb): K1, K2, K2 + SGD, K2 + Adam and the training step
at a few (n, D).  One JSON line per (n, D, ring_kb).   python tools/exp_ring.py [--rings 0,160,128,96,64]"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.sweep_D import peak, timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="10x100000000,16x60000000,20x50000000,5x200000000")
    ap.add_argument("--rings", default="0,160,128,96,64")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from beyond_deep_ensembles_b200 import _lib, ops
    lib = _lib.get()
    dev = torch.device("cuda", 0)
    has_ring = lib.bde_tune(b"ring_kb", 0) == 0
    for shape in args.shapes.split(","):
        n, D = (int(v) for v in shape.split("x"))
        g = torch.Generator(device=dev).manual_seed(7)
        X = torch.empty(n, D, device=dev)
        G = torch.empty(n, D, device=dev)
        for i in range(n):
            X[i].normal_(0.0, 0.05 * (1 + 0.1 * i), generator=g)
            G[i].normal_(0.0, 1e-3, generator=g)
        out = torch.empty_like(X)
        buf, buf2 = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        sc = ops.SvgdScratch.allocate(n, dev)
        kw = dict(lr=1e-7, momentum=0.9, nesterov=True, weight_decay=3e-4)
        nk = ops.NextKernel(True, 0.01, 1.0, 50000.0) if 2 <= n <= ops.NEXT_KERNEL_MAX_PARTICLES else None
        for ring in [int(v) for v in args.rings.split(",")]:
            if ring and not has_ring:
                continue
            if has_ring:
                lib.bde_tune(b"ring_kb", ring)
            rec = {"n": n, "D": D, "ring_kb": ring}
            rec["k1_ms"] = timeit(lambda: ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0), args.iters)
            rec["k2_ms"] = timeit(lambda: ops.svgd_apply(X, G, out, sc), args.iters)
            rec["k2_sgd_ms"] = timeit(lambda: ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, **kw), args.iters)
            rec["k2_adam_ms"] = timeit(lambda: ops.svgd_apply_adam(X, G, sc, buf, buf2, step0=10, lr=1e-7), args.iters)
            if nk is not None:
                rec["train_sgd_ms"] = timeit(lambda: ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, next_kernel=nk, **kw),
                                             args.iters)
            print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
        if has_ring:
            lib.bde_tune(b"ring_kb", 0)
        del X, G, out, buf, buf2, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
