"""ncu target (not a benchmark): a few launches of the SVGD kernels at a chosen (n, D).

    ncu --set full --clock-control none --import-source on -k regex:'svgd' -s 4 -c 4 -o gpurun_out/prof_n20 \
        python tools/prof_svgd.py 20 100000000
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.empty(n, D, device=dev)
    G = torch.empty(n, D, device=dev)
    for i in range(n):
        X[i].normal_(0.0, 0.05 * (1 + 0.1 * i), generator=g)
        G[i].normal_(0.0, 1e-3, generator=g)
    out = torch.empty_like(X)
    buf = torch.zeros(D, device=dev)
    sc = ops.SvgdScratch.allocate(n, dev)
    nk = ops.NextKernel(True, 0.01, 1.0, 50000.0) if 2 <= n <= ops.NEXT_KERNEL_MAX_PARTICLES else None
    for _ in range(reps):   # per rep: K1(+K1b), K2, fused K2+SGD, (n <= 10) training step
        ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)
        ops.svgd_apply(X, G, out, sc)
        ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, lr=1e-6, momentum=0.9, nesterov=True, weight_decay=3e-4)
        if nk is not None:
            ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, lr=1e-6, momentum=0.9, nesterov=True,
                               weight_decay=3e-4, next_kernel=nk)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
