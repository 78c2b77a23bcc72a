#!/usr/bin/env python
"""f4 timing: BBBLinear's local-reparameterisation forward as ONE tcgen05 kernel (bde_bbb_linear_fwd) against the
reference layer's own CUDA branch (bbb_layers.py:66-79: stack x3, pow, clamp x3, softplus, baddbmm, sqrt, normal_, mul,
add) in eager PyTorch on the same GPU, at the CivilComments head shapes.  JSON lines; L2 is flushed between launches."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import time_kernel  # noqa: E402
from beyond_deep_ensembles_b200 import bbb_layers, ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def reference_forward(x, w_mu, w_rho, b_mu, b_rho):
    """The CUDA branch of the reference's BBBLinear.forward, verbatim in structure (bbb_layers.py:66-79, :78-79)."""
    std_w, std_b = F.softplus(w_rho), F.softplus(b_rho)
    batch_in = torch.stack((x, (x ** 2).clamp(min=1e-4)))
    batch_mat = torch.stack((w_mu.transpose(0, 1), (std_w.transpose(0, 1) ** 2).clamp(min=1e-4)))
    batch_add = torch.stack((b_mu.expand((x.shape[0], w_mu.shape[0])), (std_b ** 2).clamp(min=1e-4).expand((x.shape[0], w_mu.shape[0]))))
    batch_out = torch.baddbmm(batch_add, batch_in, batch_mat)
    eps = torch.empty_like(batch_out[0]).normal_(0, 1)
    return batch_out[0] + torch.sqrt(batch_out[1]) * eps


for batch, fin, fout in ((16, 768, 768), (16, 768, 2), (128, 768, 768), (16, 2048, 182)):
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(batch, fin, device=dev, generator=g)
    w_mu = 0.1 * torch.randn(fout, fin, device=dev, generator=g)
    w_rho = torch.full((fout, fin), -3.0, device=dev)
    b_mu, b_rho = torch.zeros(fout, device=dev), torch.full((fout,), -3.0, device=dev)
    leaves = [t.clone().requires_grad_(True) for t in (x, w_mu, w_rho, b_mu, b_rho)]
    rec = {"batch": batch, "in": fin, "out": fout}
    with torch.no_grad():
        rec["fused_fwd_us"] = round(1e3 * time_kernel(lambda: ops.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, seed=1, stream_id=2,
                                                                                 workspace=bbb_layers._workspace), 30, 5, flush), 2)
        rec["reference_eager_fwd_us"] = round(1e3 * time_kernel(lambda: reference_forward(x, w_mu, w_rho, b_mu, b_rho), 30, 5, flush), 2)

    def fb_fused():
        out = bbb_layers._BBBLinear.apply(*leaves, 1.0, None)
        out.sum().backward()

    def fb_ref():
        reference_forward(*leaves).sum().backward()
    rec["fused_fwd_bwd_us"] = round(1e3 * time_kernel(fb_fused, 20, 5, flush), 2)
    rec["reference_eager_fwd_bwd_us"] = round(1e3 * time_kernel(fb_ref, 20, 5, flush), 2)
    rec["fwd_speedup"] = round(rec["reference_eager_fwd_us"] / rec["fused_fwd_us"], 2)
    rec["weights_bytes"] = 8 * fin * fout
    print(json.dumps(rec), flush=True)


# ---- Rank1Linear.forward (rank1.py:50-64): fused (one launch) against the reference's expression in eager PyTorch ----------
def rank1_reference_forward(x, w, s_mu, s_rho, r_mu, r_rho, bias):
    s = s_mu + torch.empty_like(s_mu).normal_(0, 1) * F.softplus(s_rho)      # util.py:170-171, twice
    r = r_mu + torch.empty_like(r_mu).normal_(0, 1) * F.softplus(r_rho)
    out = F.linear(x * s, w) * r
    out += bias.unsqueeze(0)
    return out


for batch, fin, fout in ((16, 768, 768), (16, 768, 2), (128, 768, 768), (128, 64, 10)):
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(batch, fin, device=dev, generator=g)
    w = torch.randn(fout, fin, device=dev, generator=g) / fin ** 0.5
    s_mu, r_mu = torch.ones(fin, device=dev), torch.ones(fout, device=dev)
    s_rho, r_rho = torch.full((fin,), -3.0, device=dev), torch.full((fout,), -3.0, device=dev)
    bias = torch.zeros(fout, device=dev)
    leaves = [t.clone().requires_grad_(True) for t in (x, w, s_mu, s_rho, r_mu, r_rho, bias)]
    rec = {"layer": "Rank1Linear", "batch": batch, "in": fin, "out": fout}
    with torch.no_grad():
        rec["fused_fwd_us"] = round(1e3 * time_kernel(lambda: ops.rank1_linear_fwd(x, w, s_mu, s_rho, r_mu, r_rho, bias, seed=1,
                                                                                   stream_id_s=2, stream_id_r=3,
                                                                                   workspace=bbb_layers._workspace), 30, 5, flush), 2)
        rec["reference_eager_fwd_us"] = round(1e3 * time_kernel(lambda: rank1_reference_forward(x, w, s_mu, s_rho, r_mu, r_rho, bias), 30, 5, flush), 2)

    def fb_fused_r1():
        bbb_layers._Rank1Linear.apply(*leaves, None, None, 2, 3).sum().backward()

    def fb_ref_r1():
        rank1_reference_forward(*leaves).sum().backward()
    rec["fused_fwd_bwd_us"] = round(1e3 * time_kernel(fb_fused_r1, 20, 5, flush), 2)
    rec["reference_eager_fwd_bwd_us"] = round(1e3 * time_kernel(fb_ref_r1, 20, 5, flush), 2)
    rec["fwd_speedup"] = round(rec["reference_eager_fwd_us"] / rec["fused_fwd_us"], 2)
    rec["weights_bytes"] = 4 * fin * fout
    print(json.dumps(rec), flush=True)
