#!/usr/bin/env python
"""Batched SWAG sampler (16 draws, ResNet-50 size, K = 10): fast kernel against the general one (bde_tune swag_batch=1)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import time_kernel
from beyond_deep_ensembles_b200 import _lib, ops
lib = _lib.get()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
D, K, S = 23_880_960, 10, 16
mean = torch.randn(D, device=dev, generator=g) * 0.05
sq = mean * mean + 1e-4
ring = torch.randn(K, D, device=dev, generator=g) * 0.01
outs = torch.empty(S, D, device=dev)
res = {}
for name, knob in (("general", 1), ("fast", 0)):
    lib.bde_tune(b"swag_batch", knob)
    ms = time_kernel(lambda: ops.swag_sample_batch(mean, sq, ring, 3, outs, seed=1, stream_id=2), 10, 3)
    res[name] = {"ms": round(ms, 4), "GBps": round(4 * (K + 2 + S) * D / ms / 1e6)}
    res[name + "_sum"] = float(outs.double().sum())
lib.bde_tune(b"swag_batch", 0)
print(json.dumps(res))
