"""BBBLinear's local-reparameterisation forward (src/algos/bbb_layers.py:61-88) as one tcgen05 kernel
(csrc/bbb_linear.cu) against the oracle's restatement of the reference layer, on injected noise.

Tolerance: fp32 rtol 1e-5 / atol 1e-6 against the fp64 evaluation of the same formula — the kernel multiplies on the
tensor cores in 3xTF32 (hi*hi + hi*lo + lo*hi), so this is also the test that the split keeps fp32-level accuracy."""
from __future__ import annotations

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ATOL, RTOL
from oracle import bde_oracle as O


def _case(batch, fin, fout, seed, x_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = x_scale * torch.randn(batch, fin, generator=g)
    x[0, : min(3, fin)] = torch.tensor([0.0, 5e-3, -1e-2])[: min(3, fin)]     # below / at the clamp(x^2, 1e-4) edge
    w_mu = 0.1 * torch.randn(fout, fin, generator=g)
    w_rho = -3.0 + 0.5 * torch.randn(fout, fin, generator=g)
    w_rho[0, : min(2, fin)] = torch.tensor([-9.0, 25.0])[: min(2, fin)]        # sigma^2 below the clamp; softplus threshold
    b_mu = 0.1 * torch.randn(fout, generator=g)
    b_rho = -3.0 + 0.5 * torch.randn(fout, generator=g)
    eps = torch.randn(batch, fout, generator=g)
    return x, w_mu, w_rho, b_mu, b_rho, eps


SHAPES = [(16, 768, 768), (16, 768, 2), (5, 8, 50), (32, 50, 1), (130, 64, 10), (33, 100, 36), (1, 4, 3), (64, 256, 130),
          (128, 96, 257)]


@pytest.mark.gpu
@pytest.mark.parametrize("batch,fin,fout", SHAPES)
def test_bbb_linear_fwd_vs_oracle(cuda_lib, batch, fin, fout):
    """Civil head shapes (768 x 768, 768 x 2, batch 16), the UCI layers (8 -> 50 -> 1: k ragged against the 32-wide
    k block), several batch tiles (batch > 128), out-feature tiles (> 128) and ragged everything."""
    from beyond_deep_ensembles_b200 import ops
    if fin % 4:
        pytest.skip("in_features % 4 != 0 stays on the reference's forward")
    x, w_mu, w_rho, b_mu, b_rho, eps = _case(batch, fin, fout, seed=batch * 131 + fin * 7 + fout)
    d = [t.cuda() for t in (x, w_mu, w_rho, b_mu, b_rho, eps)]
    for mc in (1.0, 3.0):
        out, std, used = ops.bbb_linear_fwd(*d[:5], eps=d[5], mc_sample=mc)
        ref_out, ref_std = O.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, eps, mc, dtype=torch.float64)
        np.testing.assert_allclose(std.cpu().numpy(), ref_std.numpy(), rtol=RTOL, atol=ATOL)
        # the activation mean alone (eps = 0) and the full output.  out = mean + std * eps is a sum of two terms of
        # either sign, so its absolute tolerance scales with the size of the addends, not of the (possibly cancelled) sum
        mean, _, _ = ops.bbb_linear_fwd(*d[:5], eps=torch.zeros_like(d[5]), mc_sample=1.0)
        ref_mean, _ = O.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, torch.zeros_like(eps), 1.0, dtype=torch.float64)
        scale = (x.abs().double() @ w_mu.abs().double().t() + b_mu.abs().double())       # sum |x_k w_k|: the dot product's condition
        assert bool(((mean.cpu().double() - ref_mean).abs() <= RTOL * ref_mean.abs() + 1e-7 * scale + ATOL * 0.1).all())
        addends = (ref_mean.abs() + (ref_std * eps.double()).abs()) / mc
        err = (out.cpu().double() - ref_out).abs()
        assert bool((err <= RTOL * addends + ATOL).all()), float((err / (addends + ATOL / RTOL)).max())
        # and in the same class as the reference's own fp32 arithmetic.  (The tensor core adds its 8 products to the TMEM
        # accumulator with truncation rather than round-to-nearest, so next to a dominant term — the sigma = 25 weight
        # planted by _case contributes 600 to one variance — the small terms of the same k block lose a few more bits
        # than in an IEEE fp32 dot product: measured up to 7 ulps of the output against 1-3 for eager fp32.)
        ref32, _ = O.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, eps, mc, dtype=torch.float32)
        assert float(err.max()) <= 10.0 * float((ref32.double() - ref_out).abs().max()) + 1e-6
        assert torch.equal(used, d[5])
    # deterministic (fixed-order split-K sum) and workspace left clean: a second launch gives the same bits
    out2, _, _ = ops.bbb_linear_fwd(*d[:5], eps=d[5], mc_sample=3.0)
    assert torch.equal(out, out2)
    # without bias
    out_nb, _, _ = ops.bbb_linear_fwd(d[0], d[1], d[2], None, None, eps=d[5])
    ref_nb, _ = O.bbb_linear_fwd(x, w_mu, w_rho, None, None, eps, 1.0, dtype=torch.float64)
    mean_nb, std_nb = O.bbb_linear_fwd(x, w_mu, w_rho, None, None, torch.zeros_like(eps), 1.0, dtype=torch.float64)
    addends = mean_nb.abs() + (std_nb * eps.double()).abs()
    assert bool(((out_nb.cpu().double() - ref_nb).abs() <= RTOL * addends + ATOL).all())


@pytest.mark.gpu
def test_bbb_linear_philox_noise_is_standard_normal_and_reproducible(cuda_lib):
    from beyond_deep_ensembles_b200 import ops
    x, w_mu, w_rho, b_mu, b_rho, _ = _case(64, 64, 512, seed=3)
    d = [t.cuda() for t in (x, w_mu, w_rho, b_mu, b_rho)]
    out, std, used = ops.bbb_linear_fwd(*d, seed=17, stream_id=5)
    out2, _, used2 = ops.bbb_linear_fwd(*d, seed=17, stream_id=5)
    assert torch.equal(out, out2) and torch.equal(used, used2)
    _, _, other = ops.bbb_linear_fwd(*d, seed=17, stream_id=6)
    assert not torch.equal(used, other)
    z = O.philox_normal(64 * 512, 17, 5)          # the library's Philox stream, counter = element index / 4
    np.testing.assert_allclose(used.cpu().numpy().reshape(-1), z, rtol=1e-4, atol=1e-4)
    full, _ = O.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, used.cpu(), 1.0, dtype=torch.float64)
    m0, s0 = O.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, torch.zeros(64, 512), 1.0, dtype=torch.float64)
    assert bool(((out.cpu().double() - full).abs() <= RTOL * (m0.abs() + (s0 * used.cpu().double()).abs()) + ATOL).all())


@pytest.mark.gpu
def test_bbb_linear_backward_matches_autograd_of_the_reference_formula(cuda_lib):
    """The autograd Function's backward (plain GEMMs) against autograd through the reference's own expression."""
    from beyond_deep_ensembles_b200.bbb_layers import _BBBLinear
    x, w_mu, w_rho, b_mu, b_rho, eps = _case(16, 96, 40, seed=9)
    x[0, 2] = 0.3            # no value exactly ON a clamp edge: fp32 and fp64 would take different sides of it
    leaves = [t.cuda().requires_grad_(True) for t in (x, w_mu, w_rho, b_mu, b_rho)]
    e = eps.cuda()
    out = _BBBLinear.apply(*leaves, 2.0, e)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).cuda()
    grads = torch.autograd.grad(out, leaves, g)
    ref_leaves = [t.detach().double().requires_grad_(True) for t in leaves]
    rx, rwm, rwr, rbm, rbr = ref_leaves
    mean = F.linear(rx, rwm, rbm)
    var = F.linear((rx ** 2).clamp(min=1e-4), (F.softplus(rwr) ** 2).clamp(min=1e-4), (F.softplus(rbr) ** 2).clamp(min=1e-4))
    ref_out = (mean + torch.sqrt(var) * e.double()) / 2.0
    ref_grads = torch.autograd.grad(ref_out, ref_leaves, g.double())
    for got, ref, name in zip(grads, ref_grads, ("x", "w_mu", "w_rho", "b_mu", "b_rho")):
        scale = float(ref.abs().max())
        np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6 * max(scale, 1.0), err_msg=name)


def test_bbb_linear_host_logic_on_the_abi_double(monkeypatch):
    """CPU: install-style patched forward over the oracle-backed ABI double — training / eval noise shapes, the KL
    attribute, fall-back to the layer's own forward for inputs the kernel does not take."""
    import fake_abi
    fake = fake_abi.install(monkeypatch)
    from beyond_deep_ensembles_b200 import bbb, bbb_layers, noise, util

    class Layer(torch.nn.Module):     # the attributes of the reference's BBBLinear (bbb_layers.py:11-27)
        def __init__(self, fin, fout):
            super().__init__()
            self.sampling, self.mc_sample, self.freeze_on_eval, self.kl_on_eval, self.use_bias = "activations", 1, True, False, True
            self.in_features, self.out_features = fin, fout
            self.weight_prior = self.bias_prior = bbb.GaussianPrior(0.0, 1.0)
            self.weight, self.bias = util.GaussianParameter((fout, fin)), util.GaussianParameter((fout,))
            self.weight.blundell_init()
            self.bias.blundell_init()
            self.kl = 0

        def forward(self, x):
            raise AssertionError("the reference forward must not be reached for a supported input")

    monkeypatch.setattr(bbb_layers, "fused_forward_applies", lambda layer, inp: inp.dim() == 2)
    layer = Layer(8, 5)
    fwd = bbb_layers.make_patched_forward(Layer.forward)
    x = torch.randn(4, 8)
    tape = iter([torch.arange(20, dtype=torch.float32) / 10, torch.ones(5)])
    with noise.inject(lambda kind, numel: next(tape)):
        y = fwd(layer, x)                      # training: one draw per activation
        ref, _ = O.bbb_linear_fwd(x, layer.weight.mean, layer.weight.rho, layer.bias.mean, layer.bias.rho,
                                  (torch.arange(20, dtype=torch.float32) / 10).view(4, 5))
        np.testing.assert_allclose(y.detach().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)
        assert torch.is_tensor(layer.kl) and layer.kl.requires_grad
        y.sum().backward()
        assert layer.weight.rho.grad is not None and layer.bias.mean.grad is not None
        layer.eval()
        y_eval = fwd(layer, x)                 # eval + freeze_on_eval: one noise vector shared by the batch
        ref_eval, _ = O.bbb_linear_fwd(x, layer.weight.mean, layer.weight.rho, layer.bias.mean, layer.bias.rho, torch.ones(4, 5))
        np.testing.assert_allclose(y_eval.detach().numpy(), ref_eval.detach().numpy(), rtol=1e-6, atol=1e-7)
    assert fake.calls.count("bbb_linear") == 2
    with pytest.raises(AssertionError):
        fwd(layer, torch.randn(2, 3, 8))       # not [batch, in]: the layer's own forward
