#!/usr/bin/env python
"""C5 sweep: SVGD posterior update over n particles x D columns on ONE GPU,
D = 10 M ... 1 B, device-resident synthetic X / G.  One JSON line per (n, D) on stdout:
K1(+K1b), K2, the two-launch step, and the one-launch training step (K2 + n SGD steps + next K1/K1b).

    python tools/sweep_D.py [--n 5,10,20] [--D 10000000,...] [--iters 10]

Times are CUDA-event means over `iters` back-to-back launches after 3 warm-ups; every working set is
larger than the 126 MB L2.  GB/s are ALGORITHMIC bytes (K1 4nD, K2 12nD, step 16nD, training step
(12n+12)D) over the measured time; `frac` is against MEASURED_PEAKS.json's copy bandwidth.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

L2_REG, KGS, NDATA = 0.01, 1.0, 50000.0


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"])
    return 6650.0


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="5,10,20")
    ap.add_argument("--D", default="10000000,30000000,100000000,300000000,1000000000")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--max-gb", type=float, default=150.0, help="skip points whose X+G+out+state exceed this")
    args = ap.parse_args()
    from beyond_deep_ensembles_b200 import _lib, ops
    _lib.get()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pk = peak()
    for n in [int(v) for v in args.n.split(",")]:
        for D in [int(v) for v in args.D.split(",")]:
            need_gb = (3 * n + 2) * D * 4 / 1e9
            if need_gb > args.max_gb:
                print(json.dumps({"n": n, "D": D, "skipped": f"needs {need_gb:.0f} GB"}), flush=True)
                continue
            g = torch.Generator(device=dev).manual_seed(7)
            X = torch.empty(n, D, device=dev)
            G = torch.empty(n, D, device=dev)
            for i in range(n):  # row by row: no [n, D] temporaries
                X[i].normal_(0.0, 0.05 * (1 + 0.1 * i), generator=g)
                G[i].normal_(0.0, 1e-3, generator=g)
            out = torch.empty_like(X)
            sc = ops.SvgdScratch.allocate(n, dev)
            it = args.iters if D <= 300_000_000 else max(3, args.iters // 2)
            k1 = timeit(lambda: ops.svgd_pairdist_bandwidth(X, sc, L2_REG, KGS, NDATA), it)
            k2 = timeit(lambda: ops.svgd_apply(X, G, out, sc), it)
            st = timeit(lambda: ops.svgd_step(X, G, out, sc, L2_REG, KGS, NDATA), it)
            rec = {"n": n, "D": D,
                   "k1_ms": k1, "k1_GBps": 4 * n * D / k1 / 1e6, "k1_frac": 4 * n * D / k1 / 1e6 / pk,
                   "k2_ms": k2, "k2_GBps": 12 * n * D / k2 / 1e6, "k2_frac": 12 * n * D / k2 / 1e6 / pk,
                   "step_ms": st, "step_GBps": 16 * n * D / st / 1e6, "step_frac": 16 * n * D / st / 1e6 / pk}
            del out
            torch.cuda.empty_cache()
            # training step: K2 + n SGD(momentum, nesterov, wd) steps (+ next K1/K1b for n <= 10), X in place
            buf = torch.zeros(D, device=dev)
            kw = dict(lr=1e-6, momentum=0.9, nesterov=True, weight_decay=3e-4)
            nk = ops.NextKernel(True, L2_REG, KGS, NDATA) if 2 <= n <= ops.NEXT_KERNEL_MAX_PARTICLES else None
            if nk is None:
                def train():
                    ops.svgd_pairdist_bandwidth(X, sc, L2_REG, KGS, NDATA)
                    ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, **kw)
            else:
                def train():
                    ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, next_kernel=nk, **kw)
            tr = timeit(train, it)
            moved = (12 * n + 8) * D + (0 if nk is not None else 4 * n * D)
            rec.update(train_step_ms=tr, train_step_launches=1 if nk is not None else 2,
                       train_step_GBps=moved / tr / 1e6, train_step_frac=moved / tr / 1e6 / pk)
            print(json.dumps(rec), flush=True)
            del X, G, buf, sc
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
