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


# ---- Rank1Linear (src/algos/rank1.py:50-64): sample s / r, linear(x * s) * r + bias ----------------------------------------

def _rank1_case(batch, fin, fout, seed, bias=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, fin, generator=g)
    w = torch.randn(fout, fin, generator=g) / fin ** 0.5
    s_mu = torch.where(torch.rand(fin, generator=g) < 0.5, -1.0, 1.0)         # GaussianParameter.sign_init (util.py:163-166)
    r_mu = torch.where(torch.rand(fout, generator=g) < 0.5, -1.0, 1.0)
    s_rho = -3.0 + 0.5 * torch.randn(fin, generator=g)
    r_rho = -3.0 + 0.5 * torch.randn(fout, generator=g)
    s_rho[0], r_rho[0] = 25.0, -30.0                                            # softplus threshold / underflow
    b = 0.1 * torch.randn(fout, generator=g) if bias else None
    return x, w, s_mu, s_rho, r_mu, r_rho, b, torch.randn(fin, generator=g), torch.randn(fout, generator=g)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,fin,fout", [s for s in SHAPES if s[1] % 4 == 0] + [(16, 64, 10), (256, 1024, 512)])
def test_rank1_linear_fwd_vs_oracle(cuda_lib, batch, fin, fout):
    from beyond_deep_ensembles_b200 import ops
    for bias in (True, False):
        case = _rank1_case(batch, fin, fout, seed=batch + 3 * fin + 7 * fout, bias=bias)
        x, w, s_mu, s_rho, r_mu, r_rho, b, es, er = case
        d = [t.cuda() if t is not None else None for t in case]
        out, lin, s, r, es_used, er_used = ops.rank1_linear_fwd(*d[:7], eps_s=d[7], eps_r=d[8])
        # the two samples: the oracle's statement within fp32 tolerance (softplus is a libm call on either side), and
        # bit-identical to what GaussianParameter.sample's own kernel (K8) gives for the same noise
        _, _, s_ref, r_ref = O.rank1_linear_fwd(x, w, s_mu, s_rho, r_mu, r_rho, b, es, er, dtype=torch.float64)
        np.testing.assert_allclose(s.cpu().numpy(), s_ref.numpy(), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(r.cpu().numpy(), r_ref.numpy(), rtol=RTOL, atol=ATOL)
        s_k8, r_k8 = torch.empty_like(d[2]), torch.empty_like(d[4])
        ops.gauss_sample_fwd(d[2], d[3], s_k8, eps=d[7], seed=0, stream_id=0)
        ops.gauss_sample_fwd(d[4], d[5], r_k8, eps=d[8], seed=0, stream_id=0)
        assert torch.equal(s, s_k8) and torch.equal(r, r_k8)
        assert torch.equal(es_used.cpu(), es) and torch.equal(er_used.cpu(), er)
        s32, r32 = s.cpu(), r.cpu()
        # the product against the fp64 evaluation of the same formula on the SAME fp32 operand x * s; tolerance of an
        # fp32 dot product: rtol on the result plus 1e-7 of the sum of |terms| (the dot product's condition)
        xs = x * s32
        ref_lin = xs.double() @ w.double().t()
        scale = xs.abs().double() @ w.abs().double().t()
        err = (lin.cpu().double() - ref_lin).abs()
        assert bool((err <= RTOL * ref_lin.abs() + 1e-7 * scale + ATOL * 0.1).all()), float(err.max())
        lin32 = xs @ w.t()
        assert float(err.max()) <= 10.0 * float((lin32.double() - ref_lin).abs().max()) + 1e-6
        # epilogue: elementwise on the kernel's own lin, bit-exact
        exp = lin.cpu() * r32
        if bias:
            exp = exp + b.unsqueeze(0)
        assert torch.equal(out.cpu(), exp)
        out2 = ops.rank1_linear_fwd(*d[:7], eps_s=d[7], eps_r=d[8])[0]
        assert torch.equal(out, out2)                      # deterministic split-K, workspace left clean


@pytest.mark.gpu
def test_rank1_linear_philox_draws_are_those_of_gauss_sample(cuda_lib):
    """Without injected noise the fused layer draws s and r from the same Philox streams GaussianParameter.sample (K8)
    would use, so switching the fusion on or off does not change a seeded run."""
    from beyond_deep_ensembles_b200 import ops
    x, w, s_mu, s_rho, r_mu, r_rho, b, _, _ = [t.cuda() for t in _rank1_case(8, 96, 40, seed=4)]
    out, lin, s, r, es, er = ops.rank1_linear_fwd(x, w, s_mu, s_rho, r_mu, r_rho, b, seed=11, stream_id_s=3, stream_id_r=4)
    s_k8, r_k8 = torch.empty_like(s_mu), torch.empty_like(r_mu)
    ops.gauss_sample_fwd(s_mu, s_rho, s_k8, eps=None, seed=11, stream_id=3)
    ops.gauss_sample_fwd(r_mu, r_rho, r_k8, eps=None, seed=11, stream_id=4)
    assert torch.equal(s, s_k8) and torch.equal(r, r_k8)
    np.testing.assert_allclose(es.cpu().numpy(), O.philox_normal(96, 11, 3), rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_rank1_linear_backward_matches_autograd_of_the_reference_formula(cuda_lib):
    from beyond_deep_ensembles_b200.bbb_layers import _Rank1Linear
    case = _rank1_case(16, 96, 40, seed=9)
    leaves = [t.cuda().requires_grad_(True) for t in case[:7]]
    es, er = case[7].cuda(), case[8].cuda()
    out = _Rank1Linear.apply(*leaves, es, er, 0, 0)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).cuda()
    grads = torch.autograd.grad(out, leaves, g)
    ref_leaves = [t.detach().double().requires_grad_(True) for t in leaves]
    ref_out, _, _, _ = O.rank1_linear_fwd(*ref_leaves, es.double(), er.double(), dtype=torch.float64)
    ref_grads = torch.autograd.grad(ref_out, ref_leaves, g.double())
    for got, ref, name in zip(grads, ref_grads, ("x", "weight", "s_mu", "s_rho", "r_mu", "r_rho", "bias")):
        scale = float(ref.abs().max())
        np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6 * max(scale, 1.0), err_msg=name)


def test_rank1_linear_host_logic_on_the_abi_double(monkeypatch):
    """CPU: the patched Rank1Linear.forward over the oracle-backed ABI double — draw order (s, then r), the component
    counter and its bias row, gradients reaching every parameter, fall-back for inputs the kernel does not take."""
    import fake_abi
    fake = fake_abi.install(monkeypatch)
    from beyond_deep_ensembles_b200 import bbb_layers, noise, util

    class Layer(torch.nn.Module):     # the attributes of the reference's Rank1Linear (rank1.py:10-34)
        def __init__(self, fin, fout, components):
            super().__init__()
            self.in_features, self.out_features, self.components = fin, fout, components
            self.layer = torch.nn.Linear(fin, fout, bias=False)
            self.s = torch.nn.ModuleList([util.GaussianParameter(fin) for _ in range(components)])
            self.r = torch.nn.ModuleList([util.GaussianParameter(fout) for _ in range(components)])
            for p in list(self.s) + list(self.r):
                p.sign_init()
            self.bias = torch.nn.Parameter(torch.randn(components, fout))
            self.component_counter = 0

        def forward(self, x):
            raise AssertionError("the reference forward must not be reached for a supported input")

    monkeypatch.setattr(bbb_layers, "rank1_forward_applies", lambda layer, inp: inp.dim() == 2)
    layer = Layer(8, 5, components=2)
    fwd = bbb_layers.make_patched_rank1_forward(Layer.forward)
    x = torch.randn(4, 8)
    draws = [torch.randn(8), torch.randn(5), torch.randn(8), torch.randn(5)]
    tape = iter(draws)
    kinds = []

    def inj(kind, numel):
        kinds.append((kind, numel))
        return next(tape)

    with noise.inject(inj):
        for c in range(2):
            y = fwd(layer, x)
            ref, _, _, _ = O.rank1_linear_fwd(x, layer.layer.weight, layer.s[c].mean, layer.s[c].rho, layer.r[c].mean,
                                              layer.r[c].rho, layer.bias[c], draws[2 * c], draws[2 * c + 1])
            np.testing.assert_allclose(y.detach().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)
            assert layer.component_counter == (c + 1) % 2
        y.sum().backward()
    assert kinds == [("gauss", 8), ("gauss", 5)] * 2
    assert layer.bias.grad is not None and bool((layer.bias.grad[0] == 0).all()) and bool((layer.bias.grad[1] == 4).all())
    for p in (layer.layer.weight, layer.s[1].mean, layer.s[1].rho, layer.r[1].mean, layer.r[1].rho):
        assert p.grad is not None and bool(p.grad.abs().sum() > 0)
    assert layer.s[0].rho.grad is None
    assert fake.calls.count("rank1_linear") == 2
    with pytest.raises(AssertionError):
        fwd(layer, torch.randn(2, 3, 8))
