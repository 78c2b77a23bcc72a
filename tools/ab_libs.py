#!/usr/bin/env python
"""Same-process, interleaved A/B of two BUILDS of libbde_b200.so (e.g. lib/prev/ against lib/): every kernel form is
timed alternately with both libraries, `--rounds` times, and the per-library MEDIAN is reported — box-to-box and
minute-to-minute drift (a few per cent on the shared pool) cancels.

    python tools/ab_libs.py --a beyond_deep_ensembles_b200/lib/prev/libbde_b200.so --shapes 10x100000000,20x50000000
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.sweep_D import timeit  # noqa: E402


def load(path, _lib):
    h = C.CDLL(path)
    for name, argtypes in _lib.SIGNATURES.items():
        if not hasattr(h, name):
            continue
        fn = getattr(h, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    h.bde_error_string.argtypes = [C.c_int]
    h.bde_error_string.restype = C.c_char_p
    return h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--a", required=True, help="library A (the baseline build)")
    ap.add_argument("--b", default=None, help="library B (default: the in-tree build)")
    ap.add_argument("--shapes", default="10x100000000,16x60000000,20x50000000,5x200000000")
    ap.add_argument("--rounds", type=int, default=7)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tune-b", default="", help="comma list key=value applied to library B (bde_tune)")
    args = ap.parse_args()
    from beyond_deep_ensembles_b200 import _lib, ops
    libs = {"A": load(os.path.abspath(args.a), _lib), "B": load(os.path.abspath(args.b), _lib) if args.b else _lib.get()}
    for kv in filter(None, args.tune_b.split(",")):
        k, v = kv.split("=")
        assert libs["B"].bde_tune(k.encode(), int(v)) == 0, kv
    dev = torch.device("cuda", 0)
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
        forms = {
            "k1": lambda: ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0),
            "k2": lambda: ops.svgd_apply(X, G, out, sc),
            "k2_sgd": lambda: ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, **kw),
            "k2_adam": lambda: ops.svgd_apply_adam(X, G, sc, buf, buf2, step0=10, lr=1e-7),
        }
        if nk is not None:
            forms["train_sgd"] = lambda: ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, next_kernel=nk, **kw)
            forms["train_adam"] = lambda: ops.svgd_apply_adam(X, G, sc, buf, buf2, step0=10, lr=1e-7, next_kernel=nk)
        times = {f: {"A": [], "B": []} for f in forms}
        for _ in range(args.rounds):
            for f, fn in forms.items():
                for which in ("A", "B"):
                    _lib._handle = libs[which]
                    times[f][which].append(timeit(fn, args.iters, warmup=2))
        rec = {"n": n, "D": D}
        for f in forms:
            a, b = statistics.median(times[f]["A"]), statistics.median(times[f]["B"])
            rec[f] = {"A_ms": round(a, 4), "B_ms": round(b, 4), "B_over_A": round(b / a, 4)}
        print(json.dumps(rec), flush=True)
        del X, G, out, buf, buf2, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
