#!/usr/bin/env python
"""Where K1's fixed cost goes at small D: %globaltimer stamps written by the instrumented build of the library
(tools/build_timing_lib.sh -> lib/timing/, -DBDE_TAIL_TIMING) into bytes 192.. of the workspace header.

    BDE_B200_LIB=beyond_deep_ensembles_b200/lib/timing/libbde_b200.so python tools/exp_tail_timing.py

Stamps: 0 CTA 0 starts | 1 last CTA leaves the streaming loop | 2 last-arriving CTA holds the ticket | 3 grid sums done |
4 n x n matrix written | 5 K1b done | 6 K1b run a second time (warm instruction cache).  One JSON line per shape (means over the runs, L2 flushed before every launch)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops
lib = _lib.get()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, D, variant in ((20, 273_664, 2), (20, 1_048_576, 2), (16, 273_664, 2), (10, 273_664, 2), (10, 512, 2), (10, 100_000_000, 2)):
    lib.bde_tune(b"pairdist_variant", variant)   # the TMA-staged direct kernel carries the stamps
    X = torch.randn(n, D, device=dev) * 0.05
    sc = ops.SvgdScratch.allocate(n, dev)
    slots = sc.ws[24:32].view(torch.int64)   # ws elements are 8 bytes: bytes 192..255 of the header
    acc, runs = torch.zeros(6, dtype=torch.float64), 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = 0.0
    for r in range(runs + 3):
        flush.zero_()
        slots.zero_()
        e0.record()
        ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 768.0)
        e1.record()
        torch.cuda.synchronize()
        t = slots.cpu().double()
        if r >= 3:
            acc += (t[1:7] - t[0:6]) / 1e3
            ev += e0.elapsed_time(e1) * 1e3
    a = (acc / runs).tolist()
    print(json.dumps({"n": n, "D": D, "event_us": round(ev / runs, 1), "stream_us": round(a[0], 1), "wait_last_cta_us": round(a[1], 1),
                      "grid_sum_us": round(a[2], 1), "matrix_us": round(a[3], 1), "k1b_us": round(a[4], 1),
                      "k1b_again_warm_us": round(a[5], 1), "in_kernel_total_us": round(sum(a[:5]), 1)}), flush=True)
lib.bde_tune(b"pairdist_variant", 0)
