#!/usr/bin/env python
"""Whole `optimizer.step(forward_closure, backward_closure)` time with REAL model closures at the
BASELINE.json configs C1-C4 (SURVEY.md §8d, item ii).  Measurement harness, not a pytest module.

For every config it reports, on one B200:
  closures_ms   the forward+backward passes alone (the model's own cuDNN/cuBLAS work),
  step_ms       the drop-in optimizer's step() (wall clock, synchronised, which is what a
                training loop sees; the host-side Python of the optimizer is part of it),
  overhead_ms   step_ms - closures_ms = what the posterior update costs on top of the model,
  launches      C-ABI launches per step,
  eager_*       the same step with the reference's op sequence in eager PyTorch on the same GPU
                (tests-only restatement built on oracle/bde_oracle.py; kind "port").

    python tests/perf_whole_step.py [--configs C1,C2,C3,C4a,C4b] [--steps 5] > gpurun_out/whole_step.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import standin_models as S  # noqa: E402
from golden_models import make_mlp  # noqa: E402


PROFILE = 0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


KINDS = ("b200", "eager")


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def wall_ms(fn, steps, warmup, rounds=3):
    """Best of `rounds` timed loops of `steps` calls (wall clock, synchronised around each loop)."""
    for _ in range(warmup):
        fn()
    best = float("inf")
    for _ in range(rounds):
        _sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        _sync()
        best = min(best, 1e3 * (time.perf_counter() - t0) / steps)
    return best


# --------------------------------------------------------------------------------------
# eager restatements of the reference's step structure (tests only; arithmetic from the oracle)
# --------------------------------------------------------------------------------------
class EagerSVGD:
    """svgd.py:66-105: per particle fwd/bwd, cat+stack gather, rbf, matmul, slice+clone scatter and
    one shared base-optimizer step per particle."""

    def __init__(self, params, reset, base, n, dataset_size, l2_reg, scale=1.0):
        self.params, self.base, self.n = list(params), base, n
        self.N, self.l2, self.scale = dataset_size, l2_reg, scale
        self.particles = []
        for i in range(n):
            self.particles.append([p.detach().clone() for p in self.params])
            if i < n - 1:
                reset()

    def step(self, fwd, bwd):
        from oracle import bde_oracle as O
        total = 0.0
        grads = []
        for i in range(self.n):
            for p, x in zip(self.params, self.particles[i]):
                p.data = x
            self.base.zero_grad()
            loss = fwd()
            total = total + loss.detach()
            bwd(loss)
            grads.append(torch.cat([p.grad.flatten() for p in self.params]))
        with torch.no_grad():
            X = torch.stack([torch.cat([x.flatten() for x in part]) for part in self.particles])
            G = torch.stack(grads)
            new = O.svgd_step_reference_order(X, G, self.l2, self.scale, self.N)
            for i in range(self.n):
                off = 0
                for p, x in zip(self.params, self.particles[i]):
                    p.grad = new[i, off:off + x.numel()].view_as(x).clone()
                    p.data = x
                    off += x.numel()
                self.base.step()
        return total / self.n


class EagerSWAG:
    """swag.py:60-105: base step, then parameters_to_vector -> host, running moments on the host,
    roll of the [D, K] deviation matrix."""

    def __init__(self, params, base, K):
        self.params, self.base, self.K = list(params), base, K
        vec = nn.utils.parameters_to_vector(self.params).detach().cpu()
        self.mean, self.sq = vec.clone(), vec ** 2
        self.dev = torch.zeros(vec.numel(), K)
        self.updates = 0

    def step(self, fwd, bwd):
        from oracle import bde_oracle as O
        self.base.zero_grad()
        loss = fwd()
        bwd(loss)
        self.base.step()
        with torch.no_grad():
            self.updates += 1
            theta = nn.utils.parameters_to_vector(self.params).cpu()
            self.mean, self.sq, col = O.swag_update(theta, self.mean, self.sq, self.updates)
            self.dev = O.swag_roll_deviations(self.dev, col)
        return loss

    def sample_parameters(self):
        from oracle import bde_oracle as O
        dev_ = self.params[0].device
        m, s, d = self.mean.to(dev_), self.sq.to(dev_), self.dev.to(dev_)
        th = O.swag_sample(m, s, d, torch.randn(self.K, device=dev_), torch.randn(m.numel(), device=dev_))
        nn.utils.vector_to_parameters(th, self.params)


class EagerIVON:
    """ivorn.py:41-127: per-tensor sampling, gradient accumulation and update in eager ops."""

    def __init__(self, params, lr, prior_prec, N, mc, damping):
        self.params, self.lr, self.pp, self.N, self.mc, self.damp = list(params), lr, prior_prec, N, mc, damping
        self.mean = [p.detach().clone() for p in self.params]
        self.mom = [torch.zeros_like(p) for p in self.params]
        self.prec = [torch.full_like(p, prior_prec / N) for p in self.params]
        self.t = 0

    def step(self, fwd, bwd):
        from oracle import bde_oracle as O
        dsum = [None] * len(self.params)
        acc = [None] * len(self.params)
        total = None
        for _ in range(self.mc):
            for k, p in enumerate(self.params):
                th, dsum[k] = O.ivon_sample(self.mean[k], self.prec[k], dsum[k], torch.randn_like(self.prec[k]), self.N)
                p.data = th
            for p in self.params:
                p.grad = None
            loss = fwd()
            bwd(loss)
            total = loss if total is None else total + loss
            for k, p in enumerate(self.params):
                acc[k] = p.grad if acc[k] is None else acc[k].add_(p.grad)
        self.t += 1
        with torch.no_grad():
            for k in range(len(self.params)):
                self.mean[k], self.mom[k], self.prec[k] = O.ivon_update(
                    acc[k], dsum[k], self.mean[k], self.mom[k], self.prec[k], mc_samples=self.mc, step=self.t,
                    lr=self.lr, prior_prec=self.pp, n_eff=self.N, damping=self.damp)
        return total / self.mc


class EagerGauss(nn.Module):
    """util.py:151-183 in eager ops (sample + KL through autograd)."""

    def __init__(self, shape):
        super().__init__()
        self.mean = nn.Parameter(torch.randn(shape) * 0.1)
        self.rho = nn.Parameter(torch.full(shape, -3.0))

    def sample(self):
        std = F.softplus(self.rho)
        return self.mean + torch.randn_like(std) * std

    def kl(self, sp=1.0):
        std = F.softplus(self.rho)
        return (0.5 * (2 * torch.log(sp / std) - 1 + (std / sp).pow(2) + (self.mean / sp).pow(2))).sum()


class GaussLinear(nn.Module):
    """Weight-space Gaussian linear layer over a GaussianParameter-like class (weight + bias)."""

    def __init__(self, gp_factory, fin, fout):
        super().__init__()
        self.w = gp_factory((fout, fin))
        self.b = gp_factory((fout,))

    def forward(self, x):
        return F.linear(x, self.w.sample(), self.b.sample())


# --------------------------------------------------------------------------------------
def bench_config(name, build, steps, warmup):
    """build(kind) -> (step_fn, closure_fn, extra dict); kind in {"b200", "eager"}."""
    from beyond_deep_ensembles_b200 import _lib
    res = {}
    for kind in KINDS:
        torch.manual_seed(0)
        try:
            step_fn, closure_fn, extra = build(kind)
        except NotImplementedError:
            continue
        c_ms = wall_ms(closure_fn, steps, warmup)
        l0 = _lib.launch_count
        s_ms = wall_ms(step_fn, steps, warmup)
        launches = (_lib.launch_count - l0) / (3 * steps + warmup)
        if PROFILE and kind == "b200":   # where the host time of step() goes (cumulative, top entries) -> stderr
            import cProfile, io, pstats
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(steps):
                step_fn()
            _sync()
            pr.disable()
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(PROFILE)
            log(f"[whole_step] cProfile of {steps} x step() at {name}:\n" + buf.getvalue()[:6000])
        pre = "" if kind == "b200" else "eager_"
        res[pre + "closures_ms"] = c_ms
        res[pre + "step_ms"] = s_ms
        res[pre + "overhead_ms"] = s_ms - c_ms
        if kind == "b200":
            res["launches_per_step"] = launches
        for k, fn in (extra or {}).items():
            res[pre + k] = wall_ms(fn, max(steps, 10), warmup)
        del step_fn, closure_fn, extra
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    if "eager_overhead_ms" in res and res.get("overhead_ms", 0) > 0:
        res["overhead_speedup_vs_eager"] = res["eager_overhead_ms"] / res["overhead_ms"]
    log(f"[whole_step] {name}: " + ", ".join(f"{k}={v:.3f}" for k, v in res.items()))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4a,C4b,C5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--eager-only-cpu", action="store_true", help="debug: run only the eager port, on the CPU")
    ap.add_argument("--prebind", default="auto", choices=["auto", "on", "off"],
                    help="zero-copy gradient capture of SVGD / iVON (BayesianOptimizer.prebind_grads): A/B runs")
    ap.add_argument("--cprofile", type=int, default=0, help="print the top N cumulative cProfile entries of step() per config")
    ap.add_argument("--kinds", default="b200,eager", help="b200 = this package, eager = the reference's op sequence (tests-only port)")
    args = ap.parse_args()
    global PROFILE
    PROFILE = args.cprofile
    from beyond_deep_ensembles_b200.algo import BayesianOptimizer
    BayesianOptimizer.prebind_grads = {"auto": "auto", "on": True, "off": False}[args.prebind]
    dev = torch.device("cpu") if args.eager_only_cpu else torch.device("cuda", 0)
    kinds = ("eager",) if args.eager_only_cpu else tuple(args.kinds.split(","))
    out = run(args.configs.split(","), args.steps, args.warmup, kinds, dev)
    out["_meta"]["prebind_grads"] = args.prebind
    print(json.dumps(out, indent=1))


def run(configs, steps=5, warmup=2, kinds=("b200", "eager"), dev=None):
    """The configs' whole-step numbers as a dict (bench.py calls this with kinds=("b200",): only the package's own
    classes run then, nothing under oracle/ is imported)."""
    global KINDS
    KINDS = tuple(kinds)
    import beyond_deep_ensembles_b200 as bde
    from beyond_deep_ensembles_b200 import util as butil
    dev = torch.device("cuda", 0) if dev is None else dev
    args = argparse.Namespace(steps=steps, warmup=warmup)
    out = {}
    want = set(configs)

    def closures_of(model, x, y, loss_fn):
        def fwd():
            return loss_fn(model(*x) if isinstance(x, tuple) else model(x), y)

        def bwd(loss):
            loss.backward()
        return fwd, bwd

    def passes(model, fwd, bwd, count):
        def run():
            for _ in range(count):
                model.zero_grad(set_to_none=True)
                bwd(fwd())
        return run

    # ---- C1: UCI MLP, SVGD n = 10, Adam ------------------------------------------------
    if "C1" in want:
        def build(kind):
            model = make_mlp(8, 50).to(dev)
            x, y = torch.randn(32, 8, device=dev), torch.randn(32, device=dev)
            fwd, bwd = closures_of(model, x, y, lambda o, t: F.mse_loss(o.squeeze(-1), t))
            base = torch.optim.Adam(model.parameters(), lr=1e-3)
            reset = lambda: butil.reset_model_params(model)  # noqa: E731
            if kind == "b200":
                opt = bde.SVGDOptimizer(model.parameters(), reset, base, 10, 768, l2_reg=0.01)
            else:
                opt = EagerSVGD(model.parameters(), reset, base, 10, 768, 0.01)
            return (lambda: opt.step(fwd, bwd)), passes(model, fwd, bwd, 10), None
        out["C1_uci_mlp_svgd_n10"] = dict(D=501, **bench_config("C1", build, max(args.steps, 20), 3))

    # ---- C2: CIFAR ResNet-20-FRN, SVGD n = 20, SGD nesterov ------------------------------
    if "C2" in want:
        def build(kind):
            model = S.ResNet20FRN().to(dev)
            x, y = torch.randn(128, 3, 32, 32, device=dev), torch.randint(0, 10, (128,), device=dev)
            fwd, bwd = closures_of(model, x, y, F.cross_entropy)
            base = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-4)
            reset = lambda: butil.reset_model_params(model)  # noqa: E731
            if kind == "b200":
                opt = bde.SVGDOptimizer(model.parameters(), reset, base, 20, 50000, l2_reg=3e-4)
            else:
                opt = EagerSVGD(model.parameters(), reset, base, 20, 50000, 3e-4)
            return (lambda: opt.step(fwd, bwd)), passes(model, fwd, bwd, 20), None
        out["C2_cifar_resnet20_svgd_n20"] = dict(D=273610, **bench_config("C2", build, args.steps, args.warmup))

    # ---- C3: iWildCam ResNet-50, SWAG K = 10 ---------------------------------------------
    if "C3" in want:
        def build(kind):
            model = S.resnet50_fc182().to(dev)
            x, y = torch.randn(16, 3, 448, 448, device=dev), torch.randint(0, 182, (16,), device=dev)
            fwd, bwd = closures_of(model, x, y, F.cross_entropy)
            base = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
            if kind == "b200":
                opt = bde.SwagOptimizer(model.parameters(), base, update_interval=1, start_epoch=0, deviation_samples=10)
            else:
                opt = EagerSWAG(model.parameters(), base, 10)

            def plain():
                base.zero_grad()
                bwd(fwd())
                base.step()
            return (lambda: opt.step(fwd, bwd)), plain, {"sample_parameters_ms": opt.sample_parameters}
        out["C3_iwildcam_resnet50_swag_k10"] = dict(D=23880950, **bench_config("C3", build, args.steps, args.warmup))

    # ---- C4a: CivilComments DistilBERT, last-layer BBB, all layers trained ----------------
    ids = mask = labels = None
    if want & {"C4a", "C4b"}:
        ids = torch.randint(0, 30522, (16, 512), device=dev)
        mask = torch.ones(16, 512, dtype=torch.long, device=dev)
        labels = torch.randint(0, 2, (16,), device=dev)
    if "C4a" in want:
        def build(kind):
            if kind == "b200":
                def gp(shape):
                    g = butil.GaussianParameter(shape)
                    g.blundell_init()
                    return g
            else:
                gp = EagerGauss
            head = nn.Sequential(GaussLinear(gp, 768, 768), nn.ReLU(), GaussLinear(gp, 768, 2))
            model = S.DistilBertClassifier(head).to(dev)
            fwd, bwd = closures_of(model, (ids, mask), labels, F.cross_entropy)
            base = torch.optim.Adam(model.parameters(), lr=1e-5)
            if kind == "b200":
                opt = bde.BBBOptimizer(model.parameters(), base, bde.GaussianPrior(0.0, 1.0), 269038, mc_samples=2)
                step = lambda: opt.step(fwd, bwd)  # noqa: E731
            else:
                gps = [m for m in model.modules() if isinstance(m, EagerGauss)]

                def step():   # bbb.py:59-89 with train_all_layers and l2_scale = 0
                    base.zero_grad()
                    data = fwd() + fwd()
                    kl = sum(g.kl() for g in gps)
                    for p in model.body.parameters():
                        kl = kl + 0.0 * p.pow(2).sum()
                    loss = kl / 269038 + data / 2
                    loss.backward()
                    base.step()

            def plain():
                base.zero_grad()
                (fwd() + fwd()).backward()
                base.step()
            return step, plain, None
        out["C4a_civil_distilbert_bbb_lastlayer"] = dict(P=592130, **bench_config("C4a", build, args.steps, args.warmup))

    # ---- C4b: CivilComments DistilBERT, full-model iVON, 2 MC samples ---------------------
    if "C4b" in want:
        def build(kind):
            model = S.DistilBertClassifier().to(dev)
            fwd, bwd = closures_of(model, (ids, mask), labels, F.cross_entropy)
            if kind == "b200":
                opt = bde.iVONOptimizer(model.parameters(), lr=1e-5, prior_prec=10.0, dataset_size=269038,
                                        damping=1e-3, mc_samples=2)
            else:
                opt = EagerIVON(model.parameters(), 1e-5, 10.0, 269038, 2, 1e-3)
            return (lambda: opt.step(fwd, bwd)), passes(model, fwd, bwd, 2), None
        out["C4b_civil_distilbert_ivon"] = dict(D=66955010, **bench_config("C4b", build, args.steps, args.warmup))

    # ---- C5: optimizer-only sweep point (n = 10 x D), this library vs the reference op sequence eager on the same GPU
    if "C5" in want:
        from beyond_deep_ensembles_b200 import ops
        from oracle import bde_oracle as O
        res5 = {}
        for D5 in (10_000_000, 100_000_000):
            g = torch.Generator(device=dev).manual_seed(0)
            X = torch.randn(10, D5, device=dev, generator=g)
            X *= (0.05 * (1 + 0.1 * torch.arange(10, device=dev, dtype=torch.float32))).unsqueeze(1)
            G = torch.randn(10, D5, device=dev, generator=g) * 1e-3
            outb = torch.empty_like(X)
            sc = ops.SvgdScratch.allocate(10, dev)
            ms_b = wall_ms(lambda: ops.svgd_step(X, G, outb, sc, 0.01, 1.0, 50000.0), 10, 3)
            ms_e = wall_ms(lambda: O.svgd_step_reference_order(X, G, 0.01, 1.0, 50000.0), 3, 1)
            res5[f"D{D5}"] = {"b200_ms": ms_b, "eager_reference_order_ms": ms_e, "speedup": ms_e / ms_b,
                              "b200_GBps": 16 * 10 * D5 / ms_b / 1e6, "eager_GBps_algorithmic": 16 * 10 * D5 / ms_e / 1e6}
            log(f"[whole_step] C5 D={D5}: b200 {ms_b:.3f} ms, eager reference order {ms_e:.3f} ms ({ms_e / ms_b:.1f}x)")
            del X, G, outb
            torch.cuda.empty_cache()
        out["C5_svgd_update_only_n10"] = res5

    out["_meta"] = {"gpu": torch.cuda.get_device_name(0) if torch.cuda.is_available() else "cpu", "torch": torch.__version__,
                    "timing": "wall clock per step, best of 3 timed loops, synchronised before and after each loop",
                    "kinds": list(KINDS)}
    if "eager" in KINDS:
        out["_meta"]["eager"] = "reference op sequence in eager PyTorch on the same GPU (tests-only port)"
    return out


if __name__ == "__main__":
    main()
