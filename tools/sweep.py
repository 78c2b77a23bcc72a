"""Launch-geometry sweep of the streaming kernels on one GPU (tuning aid, not a benchmark).

    python tools/sweep.py [--d 100000000] > gpurun_out/sweep.txt
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops  # noqa: E402


def timeit(fn, iters=10, warmup=3):
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
    ap.add_argument("--d", type=int, default=100_000_000)
    ap.add_argument("--n", type=int, default=10)
    args = ap.parse_args()
    n, D = args.n, args.d
    dev = torch.device("cuda", 0)
    X = torch.randn(n, D, device=dev) * 0.05
    G = torch.randn(n, D, device=dev) * 1e-3
    out = torch.empty_like(X)
    sc = ops.SvgdScratch.allocate(n, dev)
    h = _lib.get()
    res = {}
    # copy baseline measured the same way (read+write bytes)
    a, b = X[0], out[0]
    ms = timeit(lambda: b.copy_(a))
    res["torch_copy_GBps"] = 8 * D / ms / 1e6
    big = X.view(-1)
    ms = timeit(lambda: out.view(-1).copy_(big))
    res["torch_copy_big_GBps"] = 8 * n * D / ms / 1e6
    ms = timeit(lambda: big.sum())
    res["torch_sum_GBps"] = 4 * n * D / ms / 1e6
    h.bde_tune(b"pairdist_variant", 1)
    for c in (0, 2, 3):
        h.bde_tune(b"pairdist_ctas_per_sm", c)
        ms = timeit(lambda: ops.svgd_pairdist(X, sc))
        res[f"pairdist_ctas{c}"] = {"ms": ms, "GBps": 4 * n * D / ms / 1e6}
    h.bde_tune(b"pairdist_ctas_per_sm", 0)
    for variant in (1, 2):
        h.bde_tune(b"pairdist_variant", variant)
        ms = timeit(lambda: ops.svgd_pairdist(X, sc))
        res[f"pairdist_variant{variant}"] = {"ms": ms, "GBps": 4 * n * D / ms / 1e6}
        ms = timeit(lambda: ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0))
        res[f"pairdist_bandwidth_variant{variant}"] = {"ms": ms, "GBps": 4 * n * D / ms / 1e6}
    h.bde_tune(b"pairdist_variant", 0)
    ops.svgd_bandwidth(sc, 0.01, 1.0, 50000.0)
    h.bde_tune(b"apply_variant", 1)
    for c in (0, 4, 8):
        h.bde_tune(b"apply_ctas_per_sm", c)
        ms = timeit(lambda: ops.svgd_apply(X, G, out, sc))
        res[f"apply_ctas{c}"] = {"ms": ms, "GBps": 12 * n * D / ms / 1e6}
    h.bde_tune(b"apply_ctas_per_sm", 0)
    for variant in (1, 2):
        h.bde_tune(b"apply_variant", variant)
        ms = timeit(lambda: ops.svgd_apply(X, G, out, sc))
        res[f"apply_variant{variant}"] = {"ms": ms, "GBps": 12 * n * D / ms / 1e6}
    h.bde_tune(b"apply_variant", 0)
    ms = timeit(lambda: ops.svgd_step(X, G, out, sc, 0.01, 1.0, 50000.0))
    res["svgd_step"] = {"ms": ms, "GBps": 16 * n * D / ms / 1e6}
    # elementwise family at a large D
    Dv = 64_000_000
    v = [torch.randn(Dv, device=dev) for _ in range(6)]
    v[1].abs_().add_(0.01)
    for c in (0, 4, 5, 6, 8, 32):
        h.bde_tune(b"ew_ctas_per_sm", c)
        ms = timeit(lambda: ops.swag_update(v[0], v[2], v[3], v[4], 3))
        res[f"swag_update_ctas{c}"] = {"ms": ms, "GBps": 24 * Dv / ms / 1e6}
        ms = timeit(lambda: ops.ivon_sample(v[0], v[1], v[4], v[5], n_eff=1000.0, first=False, seed=1, stream_id=1))
        res[f"ivon_sample_philox_ctas{c}"] = {"ms": ms, "GBps": 20 * Dv / ms / 1e6}
        ms = timeit(lambda: ops.ivon_sample(v[0], v[1], v[4], v[5], n_eff=1000.0, first=False, eps=v[3]))
        res[f"ivon_sample_injected_ctas{c}"] = {"ms": ms, "GBps": 24 * Dv / ms / 1e6}
        ms = timeit(lambda: ops.ivon_accumulate(v[4], v[0], first=False))
        res[f"ivon_accumulate_ctas{c}"] = {"ms": ms, "GBps": 12 * Dv / ms / 1e6}
        ms = timeit(lambda: ops.ivon_update(v[0], v[4], v[2], v[3], v[1], mc_samples=2, step=3, lr=1e-5, beta1=0.9,
                                            beta2=0.999, prior_prec=10.0, n_eff=1000.0, tempering=1.0, damping=1e-3))
        res[f"ivon_update_ctas{c}"] = {"ms": ms, "GBps": 32 * Dv / ms / 1e6}
    h.bde_tune(b"ew_ctas_per_sm", 0)
    for k, val in res.items():
        print(k, json.dumps(val))


if __name__ == "__main__":
    main()
