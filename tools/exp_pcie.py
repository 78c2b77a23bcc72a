"""Experiment: what the PCIe link of this box gives (pinned H2D, D2H, both at once) next to the end-to-end
host-buffer SVGD step at several chunk sizes.  Tuning aid for the e2e line of bench.py."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n, D = 10, 100_000_000
h = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
h2 = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
d = torch.empty((n, D), dtype=torch.float32, device=dev)
d2 = torch.empty((n, D), dtype=torch.float32, device=dev)
h.normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = wall(lambda: d.copy_(h, non_blocking=True))
print(json.dumps({"h2d_GBps": 4 * n * D / t / 1e9}))
t = wall(lambda: h2.copy_(d, non_blocking=True))
print(json.dumps({"d2h_GBps": 4 * n * D / t / 1e9}))


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


t = wall(both)
print(json.dumps({"bidirectional_each_GBps": 4 * n * D / t / 1e9}))
del d2
X = h
G = h2
G.normal_(0, 1e-3)
O = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
sc = ops.SvgdScratch.allocate(n, dev)
for chunk in (1_000_000, 2_000_000, 4_000_000, 8_000_000, 16_000_000):
    st = ops.HostStaging.allocate(n, D, chunk, dev, dX=d)
    t = wall(lambda: ops.svgd_step_host(X, G, O, st, sc, 0.01, 1.0, 50000.0), reps=2)
    print(json.dumps({"e2e_chunk_cols": chunk, "ms": t * 1e3, "GBps_algorithmic": 16 * n * D / t / 1e9,
                      "h2d_link_GBps": 8 * n * D / t / 1e9}), flush=True)
    del st
