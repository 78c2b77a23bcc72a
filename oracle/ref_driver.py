"""Drive the UNMODIFIED reference `SVGDOptimizer.step` (oracle/_ref/src/algos/svgd.py:65-105) on the synthetic
sweep workload — TEST / MEASUREMENT INFRASTRUCTURE, never imported by the product.

Users: `bench.py --impl reference` (host cores), bench.py's `cpu_baseline` and `eager_cuda` legs, tests.

The workload has no model: one flat `nn.Parameter` of D elements stands for the network's weights.  The reference
re-enters the caller's closures once per particle; here they cost nothing — `forward_closure` returns a constant
and `backward_closure` hands particle i's synthetic gradient row to `param.grad` by reference — so the time of
`step()` is the reference's own posterior-update path: `_use_particle` / `_store_grads` (clone per particle), the
gather into [n, D] (`parameters_to_vector` + `stack`), the prior term, `rbf` (`cdist`, `quantile`, `exp`, `matmul`),
`phi`, the per-particle scatter (slice + `clone`) and the n steps of the caller's base optimizer (plain SGD here,
the cheapest one: one pass over D per particle).
"""
from __future__ import annotations

import time

import torch

from . import install_ref


def load_reference():
    """Import the staged reference's svgd module (oracle/_ref must have travelled with the snapshot)."""
    install_ref.add_to_path()
    import importlib
    return importlib.import_module("src.algos.svgd")


def synth(n: int, D: int, device, seed: int = 0):
    """X_i = 0.05 (1 + 0.1 i) N(0, 1), G ~ 1e-3 N(0, 1): the sweep's distributions (bench.py uses the same)."""
    g = torch.Generator(device=device).manual_seed(seed)
    X = torch.empty(n, D, device=device)
    G = torch.empty(n, D, device=device)
    for i in range(n):
        X[i].normal_(0.0, 0.05 * (1 + 0.1 * i), generator=g)
        G[i].normal_(0.0, 1e-3, generator=g)
    return X, G


class ReferenceSvgdJob:
    """The reference's SVGDOptimizer over one flat parameter of D elements, n particles taken from X."""

    def __init__(self, X: torch.Tensor, G: torch.Tensor, l2_reg: float, kernel_grad_scale: float, dataset_size: float,
                 lr: float = 1e-6):
        ref = load_reference()
        n, D = X.shape
        self.n, self.D, self.G = n, D, G
        self.param = torch.nn.Parameter(X[0].clone())
        self.base = torch.optim.SGD([self.param], lr=lr)
        rows = iter(range(1, n))

        def reset():   # the reference calls this n - 1 times to draw the other particles (svgd.py:58-63)
            with torch.no_grad():
                self.param.copy_(X[next(rows)])

        self.opt = ref.SVGDOptimizer([self.param], reset, self.base, n, dataset_size, l2_reg=l2_reg,
                                     kernel_grad_scale=kernel_grad_scale)
        self._i = 0
        self._loss = torch.zeros((), device=X.device)

    def _forward(self):
        return self._loss

    def _backward(self, loss):
        self.param.grad = self.G[self._i % self.n]
        self._i += 1

    def step(self):
        return self.opt.step(self._forward, self._backward)

    def particles(self):
        return torch.stack([self.opt.state[self.param][f"particle_{i}"] for i in range(self.n)])


def time_reference_steps(job: ReferenceSvgdJob, steps: int, warmup: int, sync=None):
    """Mean wall-clock ms of job.step() (`sync()` brackets the timed region on a GPU)."""
    for _ in range(warmup):
        job.step()
    if sync:
        sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        job.step()
    if sync:
        sync()
    return 1e3 * (time.perf_counter() - t0) / steps
