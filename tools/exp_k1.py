#!/usr/bin/env python
"""K1 A/B at n = 16 / 20: direct TMA-staged kernel (pairdist_variant 2) against the centred-Gram kernel
(variant 3, both warp pairings).  One JSON line per (n, D, form):  python tools/exp_k1.py [--n 16,20] [--D ...]"""
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
    ap.add_argument("--n", default="16,20")
    ap.add_argument("--D", default="50000000")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from beyond_deep_ensembles_b200 import _lib, ops
    lib = _lib.get()
    dev = torch.device("cuda", 0)
    pk = peak()
    for n in [int(v) for v in args.n.split(",")]:
        for D in [int(v) for v in args.D.split(",")]:
            g = torch.Generator(device=dev).manual_seed(7)
            X = torch.empty(n, D, device=dev)
            for i in range(n):
                X[i].normal_(0.0, 0.05 * (1 + 0.1 * i), generator=g)
            sc = ops.SvgdScratch.allocate(n, dev)
            ref = None
            forms = [("direct_tma", 2, 0, 0, 0)]
            for fold, name in ((1, "producer_warpgroup_setmaxnreg"), (2, "ninth_warp_168regs")):
                forms.append((f"gram_{name}", 3, 0, fold, 0))
            forms.append(("auto", 0, 0, 0, 0))
            for form, variant, pairing, fold, promo in forms:
                lib.bde_tune(b"pairdist_variant", variant)
                lib.bde_tune(b"gram_pairing", pairing)
                lib.bde_tune(b"gram_fold", fold)
                lib.bde_tune(b"gram_l2_promotion", promo)
                ms = timeit(lambda: ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0), args.iters)
                d = sc.dist.clone()
                if ref is None:
                    ref = d
                rel = float(((d - ref).abs() / ref.clamp_min(1e-300)).max())
                print(json.dumps({"n": n, "D": D, "form": form, "k1_ms": ms, "GBps": 4 * n * D / ms / 1e6,
                                  "frac_of_measured_peak": 4 * n * D / ms / 1e6 / pk, "redo": sc.exact_redo(),
                                  "max_rel_diff_vs_direct": rel, "sel": sc.sel.cpu().tolist()}), flush=True)
            lib.bde_tune(b"pairdist_variant", 0)
            lib.bde_tune(b"gram_pairing", 0)
            lib.bde_tune(b"gram_fold", 0)
            lib.bde_tune(b"gram_l2_promotion", 0)
            del X, sc
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
