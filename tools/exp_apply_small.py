#!/usr/bin/env python
"""K2 / K2f at small D: direct-LDG kernel against the TMA-staged kernel (bde_tune apply_variant 1 / 2) — where the auto rule of
launch_apply_opt should switch.  L2 flushed before every launch.  One JSON line per shape."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops
lib = _lib.get()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / iters * 1e3, 1)


for n, D in ((20, 65_536), (20, 131_072), (20, 273_664), (20, 524_288), (16, 131_072), (16, 273_664), (10, 65_536), (10, 131_072),
             (10, 273_664), (10, 524_288), (5, 273_664)):
    X = torch.randn(n, D, device=dev) * 0.05
    G = torch.randn(n, D, device=dev) * 1e-3
    out = torch.empty_like(X)
    buf = torch.zeros(D, device=dev)
    sc = ops.SvgdScratch.allocate(n, dev)
    ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 768.0)
    kw = dict(lr=1e-7, momentum=0.9, nesterov=True, weight_decay=3e-4)
    rec = {"n": n, "D": D}
    for name, v in (("direct", 1), ("staged", 2), ("auto", 0)):
        lib.bde_tune(b"apply_variant", v)
        rec[f"k2_{name}_us"] = timeit(lambda: ops.svgd_apply(X, G, out, sc))
        rec[f"k2f_sgd_{name}_us"] = timeit(lambda: ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, **kw))
        rec[f"step_{name}_us"] = timeit(lambda: ops.svgd_step(X, G, out, sc, 0.01, 1.0, 768.0))
    lib.bde_tune(b"apply_variant", 0)
    print(json.dumps(rec), flush=True)
