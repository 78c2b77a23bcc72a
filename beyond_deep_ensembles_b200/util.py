"""Gaussian variational parameter with reparameterised sampling (reference: src/algos/util.py:151-186).

`GaussianParameter.sample()` draws w = mean + eps * softplus(rho) with ONE fused kernel (Philox
noise generated in-kernel) and a custom backward that regenerates eps instead of storing it:
  d/dmean = dL/dw,   d/drho = dL/dw * eps * sigmoid(rho).
`GaussianParameter.kl_divergence(prior)` evaluates the prior KL and its analytic gradient with
the K9 kernels for the priors of bbb.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import noise, ops


class _GaussSample(torch.autograd.Function):
    """K8: w = mu + eps * softplus(rho)."""

    @staticmethod
    def forward(ctx, mean, rho, elem0=0, seed=None):
        mu_c, rho_c = mean.detach().contiguous(), rho.detach().contiguous()
        w = torch.empty_like(mu_c)
        eps = noise.draw("gauss", mu_c.numel(), mu_c.device)
        seed, sid = (noise.seed() if seed is None else seed), noise.next_stream_id()
        ops.gauss_sample_fwd(mu_c, rho_c, w, eps=eps, seed=seed, stream_id=sid, elem0=elem0)
        ctx.save_for_backward(rho_c, eps if eps is not None else torch.empty(0, device=mu_c.device))
        ctx.philox = (seed, sid, elem0)
        ctx.injected = eps is not None
        return w.view_as(mean)

    @staticmethod
    def backward(ctx, grad_w):
        rho_c, eps = ctx.saved_tensors
        g = grad_w.contiguous()
        grad_rho = torch.empty_like(rho_c)
        seed, sid, elem0 = ctx.philox
        ops.gauss_sample_bwd(g, rho_c, grad_rho, eps=eps if ctx.injected else None, seed=seed, stream_id=sid,
                             elem0=elem0)
        return grad_w, grad_rho.view_as(grad_w), None, None


class _KLGauss(torch.autograd.Function):
    """K9: KL(N(mu, softplus(rho)^2) || N(mu_p, sigma_p^2)) and its analytic gradient."""

    @staticmethod
    def forward(ctx, mean, rho, prior_mu, prior_sigma):
        mu_c, rho_c = mean.detach().contiguous(), rho.detach().contiguous()
        value = torch.zeros((), dtype=torch.float64, device=mu_c.device)
        ops.kl_gauss(mu_c, rho_c, prior_mu, prior_sigma, value=value, ws=ops.value_workspace(mu_c.device))
        ctx.save_for_backward(mu_c, rho_c)
        ctx.prior = (prior_mu, prior_sigma)
        return value.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        mu_c, rho_c = ctx.saved_tensors
        gmu, grho = torch.empty_like(mu_c), torch.empty_like(rho_c)
        scale = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        ops.kl_gauss(mu_c, rho_c, ctx.prior[0], ctx.prior[1], grad_mu=gmu, grad_rho=grho, grad_scale_dev=scale)
        return gmu, grho, None, None


class _KLMixture(torch.autograd.Function):
    """K9b: scale-mixture prior of bbb.py:23-37 (gradient w.r.t. the mean only)."""

    @staticmethod
    def forward(ctx, mean, pi, sigma1, sigma2):
        mu_c = mean.detach().contiguous()
        value = torch.zeros((), dtype=torch.float64, device=mu_c.device)
        ops.kl_mixture(mu_c, pi, sigma1, sigma2, value=value, ws=ops.value_workspace(mu_c.device))
        ctx.save_for_backward(mu_c)
        ctx.prior = (pi, sigma1, sigma2)
        return value.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        (mu_c,) = ctx.saved_tensors
        gmu = torch.empty_like(mu_c)
        scale = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        ops.kl_mixture(mu_c, *ctx.prior, grad_mu=gmu, grad_scale_dev=scale)
        return gmu, None, None, None


class _L2(torch.autograd.Function):
    """K10: l2_scale/2 * ||theta||^2 (bbb.py:76)."""

    @staticmethod
    def forward(ctx, theta, l2_scale):
        t = theta.detach().contiguous()
        value = torch.zeros((), dtype=torch.float64, device=t.device)
        ops.l2_term(t, l2_scale, value=value, ws=ops.value_workspace(t.device))
        ctx.save_for_backward(t)
        ctx.l2_scale = l2_scale
        return value.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        (t,) = ctx.saved_tensors
        g = torch.empty_like(t)
        scale = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        ops.l2_term(t, ctx.l2_scale, grad=g, grad_scale_dev=scale)
        return g, None


class _PriorTerms(torch.autograd.Function):
    """The whole prior term of one BBB step — KL of every Gaussian parameter plus the L2 penalty of every
    deterministic parameter (bbb.py:69-76) — as ONE autograd node: one launch for the value, one for all
    gradients (written into a single flat buffer the per-tensor gradients are views of)."""

    @staticmethod
    def forward(ctx, spec, *tensors):
        kinds, l2_scales, prior, n_seg = spec
        a = [t.detach() for t in tensors[:n_seg]]
        b_iter = iter(t.detach() for t in tensors[n_seg:])
        b = [next(b_iter) if k == ops.PRIOR_GAUSS else None for k in kinds]
        dev = a[0].device
        value = torch.zeros((), dtype=torch.float64, device=dev)
        ops.prior_terms(kinds, a, b, l2_scales=l2_scales, prior=prior, value=value, ws=ops.value_workspace(dev))
        ctx.spec, ctx.a, ctx.b = spec, a, b
        return value.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        kinds, l2_scales, prior, n_seg = ctx.spec
        a, b = ctx.a, ctx.b
        # one flat gradient buffer, every tensor's slice starting on a 16-byte boundary
        offs, total = [], 0
        for t in a + [t for t in b if t is not None]:
            offs.append(total)
            total += (t.numel() + 3) // 4 * 4
        flat = torch.empty(total, dtype=torch.float32, device=a[0].device)
        views = [flat[o:o + t.numel()] for o, t in zip(offs, a + [t for t in b if t is not None])]
        ga = views[:n_seg]
        gb_iter = iter(views[n_seg:])
        gb = [next(gb_iter) if t is not None else None for t in b]
        scale = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        ops.prior_terms(kinds, a, b, l2_scales=l2_scales, prior=prior, grad_a=ga, grad_b=gb, grad_scale_dev=scale)
        grads = [g.view_as(t) for g, t in zip(ga, a)] + [g.view_as(t) for g, t in zip(gb, b) if t is not None]
        return (None, *grads)


def prior_kind(prior):
    """("gauss", (mu, sigma, 0)) / ("mixture", (pi, s1, s2)) for the priors of bbb.py, else None."""
    kind = getattr(prior, "_bde_kind", None)
    cls = type(prior).__name__
    if kind == "gauss" or (kind is None and cls == "GaussianPrior"):
        if all(isinstance(getattr(prior, x, None), (int, float)) for x in ("mu", "sigma")):
            return "gauss", (float(prior.mu), float(prior.sigma), 0.0)
    if kind == "mixture" or (kind is None and cls == "MixturePrior"):
        if hasattr(prior, "pi") and hasattr(prior, "sigma1") and hasattr(prior, "sigma2"):
            return "mixture", (float(prior.pi), float(prior.sigma1), float(prior.sigma2))
    return None


def prior_terms(gaussians, prior, deterministic) -> torch.Tensor:
    """sum_i KL(N(mean_i, softplus(rho_i)^2) || prior) + sum_j l2_j/2 ||theta_j||^2 as one autograd node.

    gaussians: list of (mean, rho); deterministic: list of (theta, l2_scale); prior: GaussianPrior /
    MixturePrior of bbb.py (required when `gaussians` is not empty)."""
    kinds, a, b, l2 = [], [], [], []
    ptuple = (0.0, 1.0, 0.0)
    if gaussians:
        pk = prior_kind(prior)
        if pk is None:
            raise ValueError("prior_terms handles the Gaussian and scale-mixture priors of bbb.py")
        code = ops.PRIOR_GAUSS if pk[0] == "gauss" else ops.PRIOR_MIXTURE
        ptuple = pk[1]
        for mean, rho in gaussians:
            kinds.append(code)
            a.append(mean)
            l2.append(0.0)
            if code == ops.PRIOR_GAUSS:
                b.append(rho)
    for theta, scale in deterministic:
        kinds.append(ops.PRIOR_L2)
        a.append(theta)
        l2.append(float(scale))
    spec = (tuple(kinds), tuple(l2), ptuple, len(a))
    return _PriorTerms.apply(spec, *a, *b)


def l2_penalty(param: torch.Tensor, l2_scale: float) -> torch.Tensor:
    return _L2.apply(param, float(l2_scale))


def gaussian_sample(mean: torch.Tensor, rho: torch.Tensor, elem0: int = 0, seed=None) -> torch.Tensor:
    """elem0 / seed: position of this (slice of a) tensor in a D-sharded job's Philox stream and the group's key
    (dist.ColumnShard); the defaults are the single-rank case."""
    return _GaussSample.apply(mean, rho, int(elem0), seed)


def gaussian_kl(mean: torch.Tensor, rho: torch.Tensor, prior) -> torch.Tensor:
    """KL of one Gaussian parameter tensor against `prior` (GaussianPrior / MixturePrior of bbb.py).
    Any other prior object is evaluated by its own kl_divergence(mean, std)."""
    kind = getattr(prior, "_bde_kind", None)
    if kind is None:  # the reference's own prior classes (after install()) carry plain attributes
        cls = type(prior).__name__
        if cls == "GaussianPrior" and all(isinstance(getattr(prior, a, None), (int, float)) for a in ("mu", "sigma")):
            kind = "gauss"
        elif cls == "MixturePrior" and hasattr(prior, "pi") and hasattr(prior, "sigma1"):
            kind = "mixture"
    if kind == "gauss":
        return _KLGauss.apply(mean, rho, float(prior.mu), float(prior.sigma))
    if kind == "mixture":
        return _KLMixture.apply(mean, float(prior.pi), float(prior.sigma1), float(prior.sigma2))
    return prior.kl_divergence(mean, F.softplus(rho))


class GaussianParameter(nn.Module):
    """Mean / rho pair with std = softplus(rho) (reference: util.py:151-183).

    The mean parameter carries `get_parameter_kl` and `_is_gaussian_mean`, the rho parameter
    `_is_gaussian_rho`, which is how BBBOptimizer tells them apart (bbb.py:73-75).
    """

    _bde_fused_kl = True  # kl_divergence is util.gaussian_kl: BBBOptimizer may batch it with the other tensors
    #: D-sharded jobs (BBBOptimizer(process_group=...)): where this rank's slice of the tensor sits in the job-wide
    #: tensor, and the Philox key the group agreed on.  Class defaults = not sharded.
    column_offset = 0
    noise_seed = None

    def __init__(self, size, device=None):
        super().__init__()
        self.overwrite_mean(torch.empty(size, device=device))
        self.rho = nn.Parameter(torch.empty(size, device=device))
        self.rho._is_gaussian_rho = True

    def blundell_init(self, mean_std=0.1):
        torch.nn.init.normal_(self.mean, 0, mean_std)
        torch.nn.init.constant_(self.rho, -3)

    def sign_init(self):
        with torch.no_grad():
            self.mean.data = (torch.rand_like(self.mean) > 0.5).float() * 2 - 1
        torch.nn.init.constant_(self.rho, -3)

    def sample(self) -> torch.Tensor:
        return gaussian_sample(self.mean, self.rho, self.column_offset, self.noise_seed)

    def kl_divergence(self, prior):
        return gaussian_kl(self.mean, self.rho, prior)

    def overwrite_mean(self, mean):
        self.mean = nn.Parameter(mean)
        self.mean.get_parameter_kl = self.kl_divergence
        self.mean._is_gaussian_mean = True

    @property
    def std(self) -> torch.Tensor:
        return F.softplus(self.rho)


def normal_like(tensor) -> torch.Tensor:
    """Standard normal noise shaped like `tensor` from the library's Philox stream (util.py:185-186)."""
    out = torch.empty_like(tensor, dtype=torch.float32).contiguous()
    injected = noise.draw("gauss", out.numel(), out.device)
    if injected is not None:
        return injected.view_as(out).to(tensor.dtype)
    ops.philox_normal(out.view(-1), noise.seed(), noise.next_stream_id())
    return out.to(tensor.dtype)


def non_mle_params(params):
    return filter(lambda p: getattr(p, "use_mle_training", False) is False, params)


def reset_model_params(model):
    """Re-initialise every submodule that implements reset_parameters (util.py:191-202)."""
    def weight_reset(m):
        reset_parameters = getattr(m, "reset_parameters", None)
        if callable(reset_parameters):
            m.reset_parameters()
    model.apply(weight_reset)
