"""Run every streaming elementwise kernel a few times at its SURVEY §8d size (ncu target, not a benchmark).

    ncu --set full --clock-control none --import-source on -k regex:'ivon|swag|l2|kl' -o gpurun_out/prof_ew python tools/prof_ew.py
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import ops  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    D, K = 23_880_960, 10
    theta = torch.randn(D, device=dev, generator=g) * 0.05
    mean = theta + torch.randn(D, device=dev, generator=g) * 0.01
    sq = mean * mean + 1e-4
    ring = torch.randn(K, D, device=dev, generator=g) * 0.01
    out = torch.empty(D, device=dev)
    outs = torch.empty(16, D, device=dev)
    for r in range(reps):
        ops.swag_update(theta, mean, sq, ring[r % K], r + 1)
        ops.swag_sample(mean, sq, ring, 3, out, seed=1, stream_id=2)
        ops.swag_sample_batch(mean, sq, ring, 3, outs, seed=1, stream_id=2)
    del theta, mean, sq, ring, out, outs
    D = 66_955_072
    mean = torch.randn(D, device=dev, generator=g) * 0.05
    prec = torch.rand(D, device=dev, generator=g) * 1e-4 + 10.0 / 269038
    mom = torch.randn(D, device=dev, generator=g) * 1e-4
    dsum = torch.randn(D, device=dev, generator=g) * 0.3
    acc = torch.randn(D, device=dev, generator=g) * 2e-5
    theta = torch.zeros(D, device=dev)
    thetas = torch.zeros(8, D, device=dev)
    grad = torch.randn(D, device=dev, generator=g) * 1e-3
    val = torch.zeros((), dtype=torch.float64, device=dev)
    ws = ops.value_workspace(dev)
    for r in range(reps):
        ops.ivon_sample(mean, prec, dsum, theta, first=False, seed=1, stream_id=3, n_eff=269038.0)
        ops.ivon_sample_batch(mean, prec, dsum, thetas, first=False, seed=1, stream_id=3, n_eff=269038.0)
        ops.ivon_accumulate(acc, grad, first=False)
        ops.ivon_update(acc, dsum, mean, mom, prec, mc_samples=2, step=100 + r, lr=1e-5, beta1=0.9, beta2=0.999,
                        prior_prec=10.0, n_eff=269038.0, tempering=1.0, damping=1e-3)
        ops.l2_term(theta, 0.01, value=val, grad=grad, accumulate=True, ws=ws)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
