#!/usr/bin/env python
"""Batched samplers at their SURVEY §8d sizes: fast kernels against the general ones (bde_tune swag_batch=1), draws per
pass, and the single-draw kernels for scale.  One JSON line per measurement.  `prof` as argv[1]: a few plain launches
only (ncu target)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import time_kernel
from beyond_deep_ensembles_b200 import _lib, ops
lib = _lib.get()
dev = torch.device("cuda", 0)
prof = len(sys.argv) > 1 and sys.argv[1] == "prof"
g = torch.Generator(device=dev).manual_seed(1)


def timed(fn):
    if prof:
        fn(); fn(); torch.cuda.synchronize()
        return 0.0
    return time_kernel(fn, 10, 3)


# ---- SWAG, ResNet-50 size, K = 10
D, K, S = 23_880_960, 10, 16
mean = torch.randn(D, device=dev, generator=g) * 0.05
sq = mean * mean + 1e-4
ring = torch.randn(K, D, device=dev, generator=g) * 0.01
outs = torch.empty(S, D, device=dev)
for name, knob, pf, sp in (("general", 1, 0, 0), ("pass16_default", 0, 0, 0), ("pass16_x1", 0, 0, 1), ("pass16_x2", 0, 0, 2),
                           ("pass16_x4", 0, 0, 4), ("pass16_x2_nopf", 0, 9, 2), ("pass16_x2_pf2", 0, 2, 2), ("pass8_x1", 8, 0, 1),
                           ("pass8_x2", 8, 0, 2)):
    if prof and name != "pass16_default":
        continue
    lib.bde_tune(b"swag_batch", knob)
    lib.bde_tune(b"batch_prefetch", pf)
    lib.bde_tune(b"batch_splits", sp)
    ms = timed(lambda: ops.swag_sample_batch(mean, sq, ring, 3, outs, seed=1, stream_id=2))
    passes = 1 if knob in (0, 1) else S // knob
    nbytes = 4 * ((K + 2) * passes + S) * D
    print(json.dumps({"kernel": "swag_sample_batch", "form": name, "ms": round(ms, 4),
                      "GBps_dram_expected": round(nbytes / max(ms, 1e-9) / 1e6), "GBps_algorithmic": round(4 * (K + 2 + S) * D / max(ms, 1e-9) / 1e6),
                      "checksum": float(outs.double().sum())}), flush=True)
lib.bde_tune(b"batch_splits", 0)
lib.bde_tune(b"swag_batch", 0)
lib.bde_tune(b"batch_prefetch", 0)
del mean, sq, ring, outs
# ---- iVON, DistilBERT size
D = 66_955_072
mean = torch.randn(D, device=dev, generator=g) * 0.05
prec = torch.rand(D, device=dev, generator=g) * 1e-4 + 10.0 / 269038
dsum = torch.zeros(D, device=dev)
theta = torch.empty(D, device=dev)
kw = dict(n_eff=269038.0, seed=1, stream_id=3)
if not prof:
    ms = timed(lambda: ops.ivon_sample(mean, prec, dsum, theta, first=False, **kw))
    print(json.dumps({"kernel": "ivon_sample", "form": "single", "ms": round(ms, 4), "GBps_algorithmic": round(20 * D / ms / 1e6)}), flush=True)
for S in (16, 8, 5):
    outs = torch.empty(S, D, device=dev)
    for name, knob, pf in (("general", 1, 0), ("fast", 0, 0), ("fast_nopf", 0, 9), ("fast_pf2", 0, 2)):
        if prof and (name != "fast" or S != 16):
            continue
        lib.bde_tune(b"swag_batch", knob)
        lib.bde_tune(b"batch_prefetch", pf)
        dsum.zero_()
        ms = timed(lambda: ops.ivon_sample_batch(mean, prec, dsum, outs, first=False, **kw))
        print(json.dumps({"kernel": "ivon_sample_batch", "S": S, "form": name, "ms": round(ms, 4),
                          "GBps_algorithmic": round(4 * (4 + S) * D / max(ms, 1e-9) / 1e6),
                          "checksum": float(outs.double().sum())}), flush=True)
    del outs
lib.bde_tune(b"swag_batch", 0)
lib.bde_tune(b"batch_prefetch", 0)
