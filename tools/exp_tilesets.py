"""Experiment (not a benchmark line): staged K2 / K2f at n > 12 — direct-LDG kernel vs the TMA ring with 3 or 4
consumer tile sets.  Prints one JSON line per (n, D, form, variant)."""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops  # noqa: E402

PEAK = 6549.1


def timed(fn, iters=10, warmup=3):
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


def main():
    lib = _lib.get()
    dev = torch.device("cuda", 0)
    cases = [(20, 50_000_000), (16, 60_000_000), (20, 5_000_000)]
    if len(sys.argv) > 2:
        cases = [(int(sys.argv[1]), int(sys.argv[2]))]
    for n, D in cases:
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn(n, D, device=dev, generator=g) * 0.05
        G = torch.randn(n, D, device=dev, generator=g) * 1e-3
        out = torch.empty_like(X)
        s0, s1 = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        sc = ops.SvgdScratch.allocate(n, dev)
        ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)
        forms = {
            "k2": (lambda: ops.svgd_apply(X, G, out, sc), 12.0 * n * D),
            "k2f_sgd": (lambda: ops.svgd_apply_sgd(X, G, sc, s0, buf_initialized=True, lr=1e-7, momentum=0.9, nesterov=True,
                                                   weight_decay=3e-4), (12.0 * n + 8.0) * D),
            "k2f_adam": (lambda: ops.svgd_apply_adam(X, G, sc, s0, s1, step0=10, lr=1e-7, beta1=0.9, beta2=0.999, eps=1e-8,
                                                     weight_decay=0.0, decoupled_weight_decay=False), (12.0 * n + 16.0) * D),
        }
        for form, (fn, nbytes) in forms.items():
            for label, variant, ts in (("direct", 1, 0), ("tma_ts3", 2, 3), ("tma_ts4", 2, 4)):
                lib.bde_tune(b"apply_variant", variant)
                lib.bde_tune(b"apply_tile_sets", ts)
                ms = timed(fn)
                lib.bde_tune(b"apply_variant", 0)
                lib.bde_tune(b"apply_tile_sets", 0)
                gbs = nbytes / (ms * 1e-3) / 1e9
                print(json.dumps({"n": n, "D": D, "form": form, "kernel": label, "ms": round(ms, 4), "GBps": round(gbs, 1),
                                  "frac_of_measured_peak": round(gbs / PEAK, 4)}), flush=True)
        del X, G, out, s0, s1
        torch.cuda.empty_cache()
    # n <= 12, every form: 1 tile set x 512 columns (4 consumer warps) vs 3 sets x 256 columns (6 warps)
    for n, D in ((10, 100_000_000), (8, 100_000_000), (5, 100_000_000)):
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn(n, D, device=dev, generator=g) * 0.05
        G = torch.randn(n, D, device=dev, generator=g) * 1e-3
        out = torch.empty_like(X)
        s0, s1 = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        sc = ops.SvgdScratch.allocate(n, dev)
        ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)
        nk = ops.NextKernel(True, 0.01, 1.0, 50000.0)
        sgd = dict(buf_initialized=True, lr=1e-7, momentum=0.9, nesterov=True, weight_decay=3e-4)
        adam = dict(step0=10, lr=1e-7, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False)
        forms = {
            "k2": (lambda: ops.svgd_apply(X, G, out, sc), 12.0 * n * D),
            "k2f_sgd": (lambda: ops.svgd_apply_sgd(X, G, sc, s0, **sgd), (12.0 * n + 8.0) * D),
            "k2f_adam": (lambda: ops.svgd_apply_adam(X, G, sc, s0, s1, **adam), (12.0 * n + 16.0) * D),
            "train_sgd": (lambda: ops.svgd_apply_sgd(X, G, sc, s0, next_kernel=nk, **sgd), (12.0 * n + 8.0) * D),
            "train_adam": (lambda: ops.svgd_apply_adam(X, G, sc, s0, s1, next_kernel=nk, **adam), (12.0 * n + 16.0) * D),
        }
        for form, (fn, nbytes) in forms.items():
            for label, ts in (("tma_1x512", 1), ("tma_3x256", 3)):
                lib.bde_tune(b"apply_variant", 2)
                lib.bde_tune(b"apply_tile_sets", ts)
                ms = timed(fn)
                lib.bde_tune(b"apply_variant", 0)
                lib.bde_tune(b"apply_tile_sets", 0)
                gbs = nbytes / (ms * 1e-3) / 1e9
                print(json.dumps({"n": n, "D": D, "form": form, "kernel": label, "ms": round(ms, 4), "GBps": round(gbs, 1),
                                  "frac_of_measured_peak": round(gbs / PEAK, 4)}), flush=True)
        del X, G, out, s0, s1
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
