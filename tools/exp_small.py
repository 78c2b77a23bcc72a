"""Small-D SVGD step: where do the microseconds go (K1+K1b vs K2, grid caps)?  Tuning aid."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops

dev = torch.device("cuda", 0)
h = _lib.get()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=30, warmup=5, cold=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        if cold:
            flush.zero_()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


for n, D in ((20, 273_664), (10, 512), (5, 512), (10, 65_536), (10, 1_048_576), (20, 1_048_576)):
    X = torch.randn(n, D, device=dev) * 0.05
    G = torch.randn(n, D, device=dev) * 1e-3
    out = torch.empty_like(X)
    sc = ops.SvgdScratch.allocate(n, dev)
    for cold in (True, False):
        for cap in (0, 1, 2):
            h.bde_tune(b"pairdist_ctas_per_sm", cap)
            t1 = timeit(lambda: ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 768.0), cold=cold)
            t1p = timeit(lambda: ops.svgd_pairdist(X, sc), cold=cold)
            print(f"n={n} D={D} cold={cold} pairdist cap={cap}: K1+K1b {t1:.1f} us, K1 only {t1p:.1f} us", flush=True)
        h.bde_tune(b"pairdist_ctas_per_sm", 0)
        tb = timeit(lambda: ops.svgd_bandwidth(sc, 0.01, 1.0, 768.0), cold=cold)
        for cap in (0, 1, 2, 4):
            h.bde_tune(b"apply_ctas_per_sm", cap)
            t2 = timeit(lambda: ops.svgd_apply(X, G, out, sc), cold=cold)
            print(f"n={n} D={D} cold={cold} apply cap={cap}: K2 {t2:.1f} us", flush=True)
        h.bde_tune(b"apply_ctas_per_sm", 0)
        ts = timeit(lambda: ops.svgd_step(X, G, out, sc, 0.01, 1.0, 768.0), cold=cold)
        print(f"n={n} D={D} cold={cold}: K1b alone {tb:.1f} us, whole step {ts:.1f} us", flush=True)
e = timeit(lambda: None, cold=False)
print(f"empty event pair {e:.1f} us")
