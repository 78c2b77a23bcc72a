"""Parity of every CUDA kernel (called through the C-ABI) against the oracle on seeded inputs,
against the reference-generated fixtures, and — at BASELINE.json's full sizes — through
size-independent properties.  Run on the B200 box:  pytest -m gpu.

Tolerances: north-star fp32 rtol 1e-5 / atol 1e-6 for fp32 outputs; the fp64 pair distances are
held to 2e-6 relative (fp32 products summed in fp64 per 128 columns); median-selection indices
must be identical to the fp64 oracle's.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL
from oracle import bde_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from beyond_deep_ensembles_b200 import ops as _ops
    return _ops


def particles(n, D, seed, ld=None, misalign=0):
    """Sweep-style data: distinct per-particle scales -> separated distances."""
    g = torch.Generator().manual_seed(seed)
    scale = 0.05 * (1 + 0.1 * torch.arange(n, dtype=torch.float32)).unsqueeze(1)
    X = scale * torch.randn(n, D, generator=g)
    G = 1e-3 * torch.randn(n, D, generator=g)
    return X, G


def dev_matrix(x, ld=None, misalign=0):
    """Copy to the GPU with row stride `ld` and an element offset (for the unaligned paths)."""
    n, D = x.shape
    ld = ld or D
    buf = torch.zeros(n * ld + misalign + 4, dtype=torch.float32, device="cuda")
    view = torch.as_strided(buf, (n, D), (ld, 1), storage_offset=misalign)
    view.copy_(x)
    return view


SVGD_CASES = [
    # n, D, ld, misalign
    (10, 501, 512, 0), (10, 501, 501, 0), (10, 4099, 4100, 0), (5, 37, 40, 0), (2, 9, 12, 0), (20, 1000, 1000, 0),
    (20, 273610, 273664, 0), (16, 257, 260, 0), (11, 1024, 1024, 0), (12, 333, 336, 0), (3, 64, 64, 0), (1, 33, 36, 0),
    (13, 700, 700, 0),        # no register-resident fast path: generic runtime-n kernels
    (32, 129, 132, 0),        # maximum particle count
    (10, 1000, 1001, 0),      # odd row stride -> scalar kernels
    (10, 1000, 1000, 1),      # misaligned base pointer -> scalar kernels
    (10, 1 << 20, 1 << 20, 0), (7, 123457, 123460, 0), (10, 3, 4, 0), (4, 1, 4, 0),
]


@pytest.mark.parametrize("n,D,ld,mis", SVGD_CASES)
def test_svgd_kernels_vs_oracle(ops, n, D, ld, mis):
    X, G = particles(n, D, seed=n * 7919 + D)
    dX, dG = dev_matrix(X, ld, mis), dev_matrix(G, ld, mis)
    dOut = dev_matrix(torch.zeros_like(X), ld, mis)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    l2, s, N = 0.01, 1.3, 50000.0

    # K1
    ops.svgd_pairdist(dX, sc)
    d_ref = O.svgd_pairdist(X)
    np.testing.assert_allclose(sc.dist.cpu().numpy(), d_ref.numpy(), rtol=2e-6, atol=1e-12)
    assert torch.equal(sc.dist, sc.dist.t()) and sc.dist.diagonal().eq(0).all()

    # K1b on the oracle's distances: identical selection indices, K/A/h to fp32/fp64 rounding
    sc.dist.copy_(d_ref)
    ops.svgd_bandwidth(sc, l2, s, N)
    bw = O.svgd_bandwidth(d_ref, l2, s, N)
    info = sc.info.cpu().numpy()
    assert tuple(sc.sel.cpu().tolist()) == bw["sel"]
    np.testing.assert_allclose(info[0], bw["h"], rtol=1e-12)
    np.testing.assert_allclose(info[1], bw["median"], rtol=1e-12)
    np.testing.assert_allclose(sc.K.cpu().numpy(), bw["K"].numpy(), rtol=2e-7, atol=1e-30)
    np.testing.assert_allclose(sc.A.cpu().numpy(), bw["A"].numpy(), rtol=2e-7, atol=1e-30)

    # K2 with those coefficients
    ops.svgd_apply(dX, dG, dOut, sc)
    ref = O.svgd_apply(X, G, sc.K.cpu(), sc.A.cpu())
    np.testing.assert_allclose(dOut.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)

    # whole step (K1 with fused K1b tail, then K2) against the fp64 oracle
    dOut.zero_()
    ops.svgd_step(dX, dG, dOut, sc, l2, s, N)
    ref_step, info_ref = O.svgd_step_fused(X, G, l2, s, N)
    assert tuple(sc.sel.cpu().tolist()) == info_ref["sel"]
    np.testing.assert_allclose(dOut.cpu().numpy(), ref_step.numpy(), rtol=RTOL, atol=ATOL)
    # and against the reference's own op order evaluated in fp64
    ro = O.svgd_step_reference_order(X, G, l2, s, N, dtype=torch.float64)
    np.testing.assert_allclose(dOut.cpu().numpy(), ro.numpy(), rtol=RTOL, atol=ATOL)


RBF_CASES = [(5, 37), (10, 501), (20, 1000), (3, 64), (2, 9), (10, 4099), (16, 257), (1, 33)]


@pytest.mark.parametrize("n,D", RBF_CASES)
def test_svgd_vs_reference_rbf_fixture(ops, golden, n, D):
    """Against rbf() of the unmodified reference (fp64 run): K and grad_kernel."""
    g = golden("rbf.npz")
    key = f"n{n}_D{D}"
    X = torch.from_numpy(g[f"{key}_X"])
    ldp = (D + 3) // 4 * 4
    dX = dev_matrix(X, ldp)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    ops.svgd_pairdist(dX, sc)
    ops.svgd_bandwidth(sc, 0.0, 1.0, 1.0)
    np.testing.assert_allclose(sc.K.cpu().numpy(), g[f"{key}_K64"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(sc.info[1].item(), g[f"{key}_median64"], rtol=1e-6)
    out = dev_matrix(torch.zeros_like(X), ldp)
    ops.svgd_apply(dX, dev_matrix(torch.zeros_like(X), ldp), out, sc)
    np.testing.assert_allclose(-out.cpu().numpy(), g[f"{key}_gK64"], rtol=RTOL, atol=ATOL * max(1.0, np.abs(g[f"{key}_gK64"]).max()))
    ops.svgd_bandwidth(sc, 0.0, 1.0, 1.0, h_override=0.7)
    np.testing.assert_allclose(sc.K.cpu().numpy(), g[f"{key}_K64_h07"], rtol=RTOL, atol=ATOL)


def test_pairdist_accumulate_equals_single_pass_and_shards(ops):
    n, D = 10, 300_000
    X, _ = particles(n, D, 5)
    dX = dev_matrix(X)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    full = ops.svgd_pairdist(dX, sc).clone()
    sc.dist.zero_()
    for lo, hi in ((0, 100_032), (100_032, 200_000), (200_000, D)):
        ops.svgd_pairdist(dX[:, lo:hi], sc, accumulate=True)
    np.testing.assert_allclose(sc.dist.cpu().numpy(), full.cpu().numpy(), rtol=1e-7)  # fp32 partials regroup
    # determinism: two launches give bit-identical fp64 results
    again = ops.svgd_pairdist(dX, sc).clone()
    assert torch.equal(again, full)


def test_apply_identity_coefficients_are_exact(ops):
    """Size-independent property: K = I, A = 0 reproduces G bit-exactly; K = 0, A = I gives X."""
    n, D = 10, 1_000_003
    X, G = particles(n, D, 9)
    ld = D + 1  # multiple of 4
    dX, dG, dOut = dev_matrix(X, ld), dev_matrix(G, ld), dev_matrix(torch.zeros_like(X), ld)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    sc.K.copy_(torch.eye(n)); sc.A.zero_()
    ops.svgd_apply(dX, dG, dOut, sc)
    assert torch.equal(dOut, dG)
    sc.A.copy_(torch.eye(n)); sc.K.zero_()
    ops.svgd_apply(dX, dG, dOut, sc)
    assert torch.equal(dOut, dX)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n,D", [(10, 4099), (10, 501), (5, 37), (20, 1000), (2, 9), (16, 2051), (12, 70_001), (3, 4), (10, 3),
                                 (10, 300_003), (20, 273_610)])
def test_apply_variants_agree_with_oracle(ops, cuda_lib, variant, n, D):
    """Both K2 implementations (direct-LDG and TMA-staged ring) on ragged shapes, forced via bde_tune."""
    X, G = particles(n, D, seed=n + D)
    ld = (D + 3) // 4 * 4
    dX, dG, dOut = dev_matrix(X, ld), dev_matrix(G, ld), dev_matrix(torch.full_like(X, float("nan")), ld)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    ops.svgd_pairdist_bandwidth(dX, sc, 0.01, 1.0, 50000.0)
    cuda_lib.bde_tune(b"apply_variant", variant)
    try:
        ops.svgd_apply(dX, dG, dOut, sc)
        ops.svgd_apply(dX, dG, dOut, sc)  # ring state must be clean across launches
    finally:
        cuda_lib.bde_tune(b"apply_variant", 0)
    ref = O.svgd_apply(X, G, sc.K.cpu(), sc.A.cpu())
    np.testing.assert_allclose(dOut.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n,D", [(10, 4099), (10, 501), (5, 37), (2, 9), (12, 70_001), (11, 5_003), (3, 4), (10, 3),
                                 (10, 2_000_003), (7, 1_048_576)])
def test_pairdist_variants_agree_with_oracle(ops, cuda_lib, variant, n, D):
    """Both K1 implementations (direct-LDG and TMA-staged ring), with and without the fused K1b tail."""
    X, _ = particles(n, D, seed=3 * n + D)
    dX = dev_matrix(X, (D + 3) // 4 * 4)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    d_ref = O.svgd_pairdist(X)
    bw = O.svgd_bandwidth(d_ref, 0.01, 1.0, 50000.0)
    cuda_lib.bde_tune(b"pairdist_variant", variant)
    try:
        ops.svgd_pairdist(dX, sc)
        first = sc.dist.clone()
        np.testing.assert_allclose(first.cpu().numpy(), d_ref.numpy(), rtol=2e-6, atol=1e-12)
        ops.svgd_pairdist(dX, sc)  # barriers / ticket must be clean across launches; deterministic
        assert torch.equal(sc.dist, first)
        ops.svgd_pairdist(dX, sc, accumulate=True)
        np.testing.assert_allclose(sc.dist.cpu().numpy(), 2 * d_ref.numpy(), rtol=2e-6, atol=1e-12)
        ops.svgd_pairdist_bandwidth(dX, sc, 0.01, 1.0, 50000.0)
    finally:
        cuda_lib.bde_tune(b"pairdist_variant", 0)
    assert tuple(sc.sel.cpu().tolist()) == bw["sel"]
    np.testing.assert_allclose(sc.K.cpu().numpy(), bw["K"].numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(sc.A.cpu().numpy(), bw["A"].numpy(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("tile_sets", [1, 3])
@pytest.mark.parametrize("n,D", [(10, 1_200_003), (8, 900_001), (12, 500_000)])
def test_staged_plain_apply_geometries(ops, cuda_lib, tile_sets, n, D):
    """Staged plain K2 at n <= 12: both consumer geometries (1 set x 512 columns, 3 sets x 256 columns — the library
    picks by n and by the row stride) against the oracle and bit for bit against the direct-LDG kernel."""
    X, G = particles(n, D, seed=5 * n + D)
    ld = (D + 3) // 4 * 4
    sc = ops.SvgdScratch.allocate(n, "cuda")
    dX, dG = dev_matrix(X, ld), dev_matrix(G, ld)
    ops.svgd_pairdist_bandwidth(dX, sc, 0.01, 1.0, 768.0)
    outs = {}
    for variant in (1, 2):
        dOut = dev_matrix(torch.full_like(X, float("nan")), ld)
        cuda_lib.bde_tune(b"apply_variant", variant)
        cuda_lib.bde_tune(b"apply_tile_sets", tile_sets)
        try:
            ops.svgd_apply(dX, dG, dOut, sc)
            ops.svgd_apply(dX, dG, dOut, sc)
        finally:
            cuda_lib.bde_tune(b"apply_variant", 0)
            cuda_lib.bde_tune(b"apply_tile_sets", 0)
        outs[variant] = dOut.cpu()
    assert torch.equal(outs[1], outs[2])
    np.testing.assert_allclose(outs[2].numpy(), O.svgd_apply(X, G, sc.K.cpu(), sc.A.cpu()).numpy(), rtol=RTOL, atol=ATOL)


def test_apply_rejects_overlap(ops):
    from beyond_deep_ensembles_b200 import _lib
    n, D = 4, 64
    X, G = particles(n, D, 1)
    dX, dG = dev_matrix(X), dev_matrix(G)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    with pytest.raises(_lib.BdeError):
        ops.svgd_apply(dX, dG, dG, sc)


@pytest.mark.parametrize("D", [100_000_000])
def test_svgd_full_size_properties(ops, D):
    """BASELINE sweep size (n=10, D=1e8 per GPU): checked through properties and a chunked
    fp64 evaluation on the device (torch ops, independent of the library)."""
    n = 10
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(n, D, device="cuda", generator=g)
    X *= (0.05 * (1 + 0.1 * torch.arange(n, device="cuda", dtype=torch.float32))).unsqueeze(1)
    G = torch.randn(n, D, device="cuda", generator=g) * 1e-3
    out = torch.empty_like(X)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    ops.svgd_step(X, G, out, sc, 0.01, 1.0, 50000.0)
    d = sc.dist.clone()
    # (1) distances vs chunked fp64 torch evaluation
    ref = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    step = 1 << 22
    for c0 in range(0, D, step):
        x = X[:, c0:c0 + step].double()
        ref += torch.cdist(x, x, p=2, compute_mode="donot_use_mm_for_euclid_dist") ** 2
    np.testing.assert_allclose(d.cpu().numpy(), ref.cpu().numpy(), rtol=2e-6)
    bw = O.svgd_bandwidth(ref.cpu(), 0.01, 1.0, 50000.0)
    assert tuple(sc.sel.cpu().tolist()) == bw["sel"]
    # (2) out on a random column sample vs the oracle with the device's K and A
    cols = torch.randint(0, D, (4096,), generator=torch.Generator().manual_seed(1))
    cols = torch.cat([cols, torch.tensor([0, 1, 2, 3, D - 4, D - 3, D - 2, D - 1])]).cuda()
    ref_cols = O.svgd_apply(X[:, cols].cpu(), G[:, cols].cpu(), bw["K"], bw["A"])
    np.testing.assert_allclose(out[:, cols].cpu().numpy(), ref_cols.numpy(), rtol=RTOL, atol=ATOL)
    # (3) linearity in G: step(X, 2G) - step(X, G) == K G (same K since X unchanged)
    out2 = torch.empty_like(X)
    G.mul_(2.0)
    ops.svgd_apply(X, G, out2, sc)
    diff = (out2[:, cols] - out[:, cols]).cpu().double()
    KG = bw["K"] @ (G[:, cols].cpu().double() * 0.5)
    np.testing.assert_allclose(diff.numpy(), KG.numpy(), rtol=2e-4, atol=1e-9)
    # (4) permutation equivariance of the distances (exact: same per-pair arithmetic order)
    perm = torch.tensor([3, 0, 9, 1, 7, 2, 8, 4, 6, 5], device="cuda")
    Xp = X[perm].contiguous()
    ops.svgd_pairdist(Xp, sc)
    np.testing.assert_allclose(sc.dist.cpu().numpy(), d[perm][:, perm].cpu().numpy(), rtol=1e-12)


# ---------------------------------------------------------------- f1: K2 fused with the base optimizer
OPT_KINDS = {
    "sgd-cifar": ("sgd", dict(lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-4)),
    "sgd-plain": ("sgd", dict(lr=0.1)),
    "sgd-damp": ("sgd", dict(lr=0.02, momentum=0.8, dampening=0.3)),
    "adam": ("adam", dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)),
    "adam-wd": ("adam", dict(lr=1e-2, betas=(0.8, 0.99), eps=1e-6, weight_decay=0.01)),
    "adamw": ("adamw", dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)),
}


def fused_apply(ops, kind, hyper, dX, dG, sc, s0, s1, initialized, step0, out_last):
    if kind == "sgd":
        ops.svgd_apply_sgd(dX, dG, sc, s0 if hyper.get("momentum", 0) != 0 else None, buf_initialized=initialized,
                           out_last=out_last, **hyper)
    else:
        ops.svgd_apply_adam(dX, dG, sc, s0, s1, step0=step0, lr=hyper["lr"], beta1=hyper["betas"][0],
                            beta2=hyper["betas"][1], eps=hyper["eps"], weight_decay=hyper["weight_decay"],
                            decoupled_weight_decay=(kind == "adamw"), out_last=out_last)


def check_update(x_new, x_old, ref_new, what):
    """Updated particles within the north-star tolerance, and the UPDATE itself (x_new - x_old, a small
    correction) to 1e-3 relative + 1e-5 of its largest entry."""
    np.testing.assert_allclose(x_new, ref_new, rtol=RTOL, atol=ATOL, err_msg=what)
    upd, upd_ref = x_new.astype(np.float64) - x_old, ref_new.astype(np.float64) - x_old
    bound = 1e-3 * np.abs(upd_ref) + 1e-5 * np.abs(upd_ref).max() + 6e-8 * np.abs(x_old)  # last term: fp32 rounding of x itself
    assert (np.abs(upd - upd_ref) <= bound).all(), what


@pytest.mark.parametrize("tile_sets", [3, 4])
@pytest.mark.parametrize("opt", [None, "sgd-cifar", "adam"])
@pytest.mark.parametrize("n,D", [(20, 600_001), (16, 700_003), (13, 450_000)])
def test_staged_apply_tile_sets(ops, cuda_lib, tile_sets, opt, n, D):
    """Staged K2 / K2f at n > 12: the consumer warps are split into tile sets that take turns on the ring
    (svgd_kernels.cuh, "Tile sets").  D is large enough that every set wraps the ring several times (the stage
    re-use / barrier-parity logic), with a ragged last tile and D % 4 != 0; both set counts; plain, SGD and Adam
    forms against the oracle, and bit-for-bit against the direct-LDG kernel (same per-column arithmetic)."""
    X, G = particles(n, D, seed=7 * n + D)
    X *= 4.0
    G *= 50.0
    ld = (D + 3) // 4 * 4
    sc = ops.SvgdScratch.allocate(n, "cuda")
    ops.svgd_pairdist_bandwidth(dev_matrix(X, ld), sc, 0.01, 1.0, 768.0)
    results = {}
    for variant in (1, 2):
        dX, dG = dev_matrix(X, ld), dev_matrix(G, ld)
        dOut = dev_matrix(torch.full_like(X, float("nan")), ld)
        s0, s1, out_last = (torch.zeros(D, device="cuda") for _ in range(3))
        cuda_lib.bde_tune(b"apply_variant", variant)
        cuda_lib.bde_tune(b"apply_tile_sets", tile_sets)
        try:
            if opt is None:
                ops.svgd_apply(dX, dG, dOut, sc)
                ops.svgd_apply(dX, dG, dOut, sc)   # ring state must be clean across launches
                results[variant] = (dOut.cpu(),)
            else:
                kind, hyper = OPT_KINDS[opt]
                fused_apply(ops, kind, hyper, dX, dG, sc, s0, s1, False, 0, out_last)
                fused_apply(ops, kind, hyper, dX, dG, sc, s0, s1, True, n, out_last)   # carried state, same K / A
                results[variant] = (dX.cpu(), s0.cpu(), s1.cpu(), out_last.cpu())
        finally:
            cuda_lib.bde_tune(b"apply_variant", 0)
            cuda_lib.bde_tune(b"apply_tile_sets", 0)
    for a, b in zip(results[1], results[2]):
        assert torch.equal(a, b)
    K, A = sc.K.cpu(), sc.A.cpu()
    if opt is None:
        np.testing.assert_allclose(results[2][0].numpy(), O.svgd_apply(X, G, K, A).numpy(), rtol=RTOL, atol=ATOL)
    else:
        kind, hyper = OPT_KINDS[opt]
        x1, state = O.svgd_base_optimizer_steps(X, O.svgd_apply(X, G, K, A).float(), kind, hyper, None)
        x2, state = O.svgd_base_optimizer_steps(x1, O.svgd_apply(x1, G, K, A).float(), kind, hyper, state)
        np.testing.assert_allclose(results[2][0].numpy(), x2.numpy(), rtol=RTOL, atol=5 * ATOL)


@pytest.mark.parametrize("variant", [1, 2], ids=["direct", "tma"])
@pytest.mark.parametrize("opt", list(OPT_KINDS))
@pytest.mark.parametrize("n,D,ld,mis", [(10, 4099, 4100, 0), (10, 501, 512, 0), (5, 37, 40, 0), (20, 1000, 1000, 0),
                                        (16, 2051, 2052, 0), (12, 70_001, 70_004, 0), (2, 9, 12, 0), (10, 3, 4, 0),
                                        (13, 700, 700, 0), (10, 1000, 1001, 0), (10, 1000, 1000, 1), (3, 200_000, 200_000, 0)])
def test_fused_apply_base_optimizer_vs_oracle(ops, cuda_lib, variant, opt, n, D, ld, mis):
    """bde_svgd_apply_sgd / _adam against the reference semantics (svgd.py:92-103: one shared torch.optim
    optimizer stepped once per particle) for two consecutive SVGD steps (first step: uninitialised momentum
    buffer / step count 0; second: carried state), both kernel forms, ragged and unaligned shapes."""
    kind, hyper = OPT_KINDS[opt]
    X, G = particles(n, D, seed=n * 31 + D)
    X *= 4.0            # O(0.2) particles
    G *= 50.0           # O(0.05) gradients: the update is well above fp32 rounding of x
    dX, dG = dev_matrix(X, ld, mis), dev_matrix(G, ld, mis)
    dOut = dev_matrix(torch.zeros_like(X), ld, mis)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    # 4-byte-misaligned state vectors in the misaligned case
    buf0 = torch.zeros(D + 8, device="cuda")
    s0 = buf0[mis:mis + D]
    s1 = torch.zeros(D + 8, device="cuda")[mis:mis + D]
    out_last = torch.zeros(D + 8, device="cuda")[mis:mis + D]
    state, x_ref = None, X.clone()
    cuda_lib.bde_tune(b"apply_variant", variant)
    try:
        for step in range(2):
            ops.svgd_pairdist(dX, sc)
            ops.svgd_bandwidth(sc, 0.01, 1.0, 768.0)
            x_old = dX.cpu()
            g_step = G * (1.0 + 0.5 * step)
            dG.copy_(g_step)
            ref_out = O.svgd_apply(x_old, g_step, sc.K.cpu(), sc.A.cpu()).float()
            # the new gradients as the plain K2 computes them: the optimizer arithmetic is checked tightly on
            # these (Adam's g / (|g| + eps) amplifies last-bit differences of near-zero gradients, which
            # says nothing about the optimizer step itself)
            ops.svgd_apply(dX, dG, dOut, sc)
            k2_out = dOut.cpu()
            np.testing.assert_allclose(k2_out.numpy(), ref_out.numpy(), rtol=RTOL, atol=ATOL)
            fused_apply(ops, kind, hyper, dX, dG, sc, s0, s1, step > 0, step * n, out_last)
            x_ref, state_next = O.svgd_base_optimizer_steps(x_old, k2_out, kind, hyper, state)
            check_update(dX.cpu().numpy(), x_old.numpy().astype(np.float64), x_ref.numpy(), f"step {step}")
            x_ref_full, _ = O.svgd_base_optimizer_steps(x_old, ref_out, kind, hyper, state)
            np.testing.assert_allclose(dX.cpu().numpy(), x_ref_full.numpy(), rtol=RTOL, atol=5 * ATOL)
            state = state_next
            assert torch.equal(out_last.cpu(), k2_out[n - 1])  # same arithmetic as the plain K2, bit for bit
            if kind == "sgd" and hyper.get("momentum", 0) != 0:
                np.testing.assert_allclose(s0.cpu().numpy(), state["momentum_buffer"].numpy(), rtol=1e-4, atol=1e-6)
            elif kind != "sgd":
                np.testing.assert_allclose(s0.cpu().numpy(), state["exp_avg"].numpy(), rtol=1e-4, atol=1e-7)
                np.testing.assert_allclose(s1.cpu().numpy(), state["exp_avg_sq"].numpy(), rtol=1e-4, atol=1e-9)
                assert float(state["step"]) == (step + 1) * n
            assert torch.equal(dG.cpu(), g_step)  # G untouched
    finally:
        cuda_lib.bde_tune(b"apply_variant", 0)


@pytest.mark.parametrize("variant", [1, 2], ids=["direct", "tma"])
@pytest.mark.parametrize("opt", ["sgd-cifar", "sgd-plain", "adam", "adamw"])
@pytest.mark.parametrize("n,D,ld,mis", [(10, 4099, 4100, 0), (10, 501, 512, 0), (5, 37, 40, 0), (2, 9, 12, 0), (10, 3, 4, 0),
                                        (7, 70_001, 70_004, 0), (10, 300_000, 300_000, 0), (3, 200_000, 200_000, 0),
                                        (9, 1000, 1001, 0), (10, 1000, 1000, 1), (12, 5003, 5004, 0), (20, 1000, 1000, 0),
                                        (13, 700, 700, 0)])
def test_train_step_next_distances(ops, cuda_lib, variant, opt, n, D, ld, mis):
    """bde_svgd_train_step_*: the particles are updated exactly as by bde_svgd_apply_* (bit for bit), and the
    same launch leaves the pair distances / K / A / selection of the UPDATED particles — what K1 + K1b
    return when run on them afterwards.  n > 10, unaligned rows: the documented two-launch form."""
    kind, hyper = OPT_KINDS[opt]
    X, G = particles(n, D, seed=n * 17 + D)
    X *= 4.0
    G *= 50.0
    dG = dev_matrix(G, ld, mis)
    dXa, dXb, dXc = dev_matrix(X, ld, mis), dev_matrix(X, ld, mis), dev_matrix(X, ld, mis)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    scb, scc, scr = (ops.SvgdScratch.allocate(n, "cuda") for _ in range(3))
    ops.svgd_pairdist_bandwidth(dXa, sc, 0.01, 1.0, 768.0)
    for t in (scb, scc):
        t.K.copy_(sc.K)
        t.A.copy_(sc.A)

    def state():
        return tuple(torch.zeros(D + 8, device="cuda")[mis:mis + D] for _ in range(3))

    cuda_lib.bde_tune(b"apply_variant", variant)
    try:
        for step in range(2):
            if step == 0:
                sa, sb, s_c = state(), state(), state()
            args = dict(initialized=step > 0, step0=step * n)
            fused_apply(ops, kind, hyper, dXa, dG, sc, sa[0], sa[1], out_last=sa[2], **args)
            nk = ops.NextKernel(True, 0.01, 1.0, 768.0)
            if kind == "sgd":
                ops.svgd_apply_sgd(dXb, dG, scb, sb[0] if hyper.get("momentum", 0) != 0 else None, buf_initialized=step > 0,
                                   out_last=sb[2], next_kernel=nk, **hyper)
                ops.svgd_apply_sgd(dXc, dG, scc, s_c[0] if hyper.get("momentum", 0) != 0 else None, buf_initialized=step > 0,
                                   out_last=s_c[2], next_kernel=ops.NextKernel(False, 0.01, 1.0, 768.0), **hyper)
            else:
                akw = dict(step0=step * n, lr=hyper["lr"], beta1=hyper["betas"][0], beta2=hyper["betas"][1], eps=hyper["eps"],
                           weight_decay=hyper["weight_decay"], decoupled_weight_decay=(kind == "adamw"))
                ops.svgd_apply_adam(dXb, dG, scb, sb[0], sb[1], out_last=sb[2], next_kernel=nk, **akw)
                ops.svgd_apply_adam(dXc, dG, scc, s_c[0], s_c[1], out_last=s_c[2],
                                    next_kernel=ops.NextKernel(False, 0.01, 1.0, 768.0), **akw)
            # (1) same update, bit for bit, as the launch without the distance pass
            assert torch.equal(dXb, dXa) and torch.equal(dXc, dXa)
            for a, b in zip(sa, sb):
                assert torch.equal(a, b)
            # (2) distances of the updated particles: fp64 oracle, and K1 run on them afterwards
            d_ref = O.svgd_pairdist(dXb.cpu())
            np.testing.assert_allclose(scb.dist.cpu().numpy(), d_ref.numpy(), rtol=2e-6, atol=1e-12)
            np.testing.assert_allclose(scc.dist.cpu().numpy(), scb.dist.cpu().numpy(), rtol=1e-12, atol=1e-15)
            ops.svgd_pairdist_bandwidth(dXb, scr, 0.01, 1.0, 768.0)
            # (two different groupings of the fp32 partial sums: both are ~1e-9 from the fp64 oracle at large D)
            np.testing.assert_allclose(scb.dist.cpu().numpy(), scr.dist.cpu().numpy(), rtol=5e-8, atol=1e-12)
            # (3) K1b of the next step ran in the tail: same selection, same coefficients
            bw = O.svgd_bandwidth(d_ref, 0.01, 1.0, 768.0)
            assert tuple(scb.sel.cpu().tolist()) == bw["sel"]
            np.testing.assert_allclose(scb.K.cpu().numpy(), bw["K"].numpy(), rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(scb.A.cpu().numpy(), bw["A"].numpy(), rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(scb.info.cpu().numpy()[0], bw["h"], rtol=1e-6)
            # (4) without fuse_bandwidth the coefficients are left alone; refresh them for the next round
            assert torch.equal(scc.K, sc.K) and torch.equal(scc.A, sc.A)
            sc.K.copy_(scb.K)
            sc.A.copy_(scb.A)
            scc.K.copy_(scb.K)
            scc.A.copy_(scb.A)
    finally:
        cuda_lib.bde_tune(b"apply_variant", 0)


@pytest.mark.parametrize("opt", ["sgd-cifar", "adam"])
@pytest.mark.parametrize("n,D", [(10, 2_000_003), (4, 1_500_001), (7, 40_000), (5, 900_001)])
def test_train_step_tile_geometries(ops, cuda_lib, opt, n, D):
    """Staged training-step kernel: 1 tile set x 512 columns and 3 sets x 256 columns (one is the default, by n) walk the ring
    differently (per-set flush counters, stage hand-over between sets) but must update the particles bit for
    bit alike and leave the same next-step distances; D wraps the ring many times per set, D % 4 != 0."""
    kind, hyper = OPT_KINDS[opt]
    X, G = particles(n, D, seed=n * 13 + D)
    X *= 4.0
    G *= 50.0
    ld = (D + 3) // 4 * 4
    sc0 = ops.SvgdScratch.allocate(n, "cuda")
    ops.svgd_pairdist_bandwidth(dev_matrix(X, ld), sc0, 0.01, 1.0, 768.0)
    runs = []
    for sets in (1, 3):
        dX, dG = dev_matrix(X, ld), dev_matrix(G, ld)
        sc = ops.SvgdScratch.allocate(n, "cuda")
        sc.K.copy_(sc0.K)
        sc.A.copy_(sc0.A)
        s0, s1, out_last = (torch.zeros(D, device="cuda") for _ in range(3))
        cuda_lib.bde_tune(b"apply_variant", 2)
        cuda_lib.bde_tune(b"apply_tile_sets", sets)
        try:
            for step in range(2):
                nk = ops.NextKernel(True, 0.01, 1.0, 768.0)
                if kind == "sgd":
                    ops.svgd_apply_sgd(dX, dG, sc, s0, buf_initialized=step > 0, out_last=out_last, next_kernel=nk, **hyper)
                else:
                    ops.svgd_apply_adam(dX, dG, sc, s0, s1, step0=step * n, lr=hyper["lr"], beta1=hyper["betas"][0],
                                        beta2=hyper["betas"][1], eps=hyper["eps"], weight_decay=hyper["weight_decay"],
                                        decoupled_weight_decay=False, out_last=out_last, next_kernel=nk)
                if step == 0:
                    first = (dX.cpu(), s0.cpu(), s1.cpu(), out_last.cpu())
        finally:
            cuda_lib.bde_tune(b"apply_variant", 0)
            cuda_lib.bde_tune(b"apply_tile_sets", 0)
        runs.append((first, dX.cpu(), s0.cpu(), s1.cpu(), out_last.cpu(), sc.dist.cpu(), sc.sel.cpu(), sc.K.cpu(), sc.A.cpu()))
    # step 1 (same K / A): bit for bit; step 2 runs on each geometry's own K / A, whose fp32 partial sums are
    # grouped differently (~1e-9 apart), so the particles agree to rounding only
    for a, b in zip(runs[0][0], runs[1][0]):
        assert torch.equal(a, b)
    for a, b in zip(runs[0][1:5], runs[1][1:5]):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=RTOL, atol=ATOL)
    for r in runs:
        d_ref = O.svgd_pairdist(r[1][:, :D])
        np.testing.assert_allclose(r[5].numpy(), d_ref.numpy(), rtol=2e-6, atol=1e-12)
        bw = O.svgd_bandwidth(d_ref, 0.01, 1.0, 768.0)
        assert tuple(r[6].tolist()) == bw["sel"]
        np.testing.assert_allclose(r[7].numpy(), bw["K"].numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(r[8].numpy(), bw["A"].numpy(), rtol=1e-5, atol=1e-9)


def test_train_step_full_size(ops):
    """n = 10 x D = 1e8: the single training-step launch against fused apply + K1 run one after the other."""
    n, D = 10, 100_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(n, D, device="cuda", generator=g)
    X *= (0.05 * (1 + 0.1 * torch.arange(n, device="cuda", dtype=torch.float32))).unsqueeze(1)
    G = torch.randn(n, D, device="cuda", generator=g) * 1e-2
    X2 = X.clone()
    sc, sc2, scr = (ops.SvgdScratch.allocate(n, "cuda") for _ in range(3))
    ops.svgd_pairdist_bandwidth(X, sc, 3e-4, 1.0, 50000.0)
    sc2.K.copy_(sc.K)
    sc2.A.copy_(sc.A)
    hyper = dict(lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-4)
    buf, buf2 = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=False, **hyper)
    ops.svgd_apply_sgd(X2, G, sc2, buf2, buf_initialized=False, next_kernel=ops.NextKernel(True, 3e-4, 1.0, 50000.0), **hyper)
    assert torch.equal(X2, X) and torch.equal(buf2, buf)
    del X2, buf2, G
    ops.svgd_pairdist_bandwidth(X, scr, 3e-4, 1.0, 50000.0)
    np.testing.assert_allclose(sc2.dist.cpu().numpy(), scr.dist.cpu().numpy(), rtol=1e-8)
    assert torch.equal(sc2.sel, scr.sel)
    np.testing.assert_allclose(sc2.K.cpu().numpy(), scr.K.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(sc2.A.cpu().numpy(), scr.A.cpu().numpy(), rtol=1e-6, atol=1e-12)
    # chunked fp64 distances on the device as the accuracy reference
    d = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    for c0 in range(0, D, 1 << 22):
        x = X[:, c0:c0 + (1 << 22)].double()
        d += (x.unsqueeze(1) - x.unsqueeze(0)).square().sum(dim=2)
    np.testing.assert_allclose(sc2.dist.cpu().numpy(), d.cpu().numpy(), rtol=2e-6)


def test_fused_apply_rejects_bad_arguments(ops):
    from beyond_deep_ensembles_b200 import _lib
    n, D = 4, 64
    X = torch.randn(n, D, device="cuda")
    sc = ops.SvgdScratch.allocate(n, "cuda")
    with pytest.raises(ValueError):
        ops.svgd_apply_sgd(X, X.clone(), sc, None, buf_initialized=False, lr=0.1, momentum=0.9)
    with pytest.raises(_lib.BdeError):  # X and G overlap
        ops.svgd_apply_sgd(X, X, sc, None, buf_initialized=False, lr=0.1)


def test_fused_apply_full_size_properties(ops):
    """n = 10 x D = 1e8 (the sweep point): lr = 0 leaves X bit-identical while the momentum buffer follows the
    reference recurrence; a real step matches the oracle on a random column sample; padding columns stay 0."""
    n, D = 10, 100_000_000
    g = torch.Generator(device="cuda").manual_seed(11)
    X = torch.randn(n, D, device="cuda", generator=g)
    X *= (0.05 * (1 + 0.1 * torch.arange(n, device="cuda", dtype=torch.float32))).unsqueeze(1)
    G = torch.randn(n, D, device="cuda", generator=g) * 1e-2
    X[:, -64:] = 0.0
    G[:, -64:] = 0.0  # arena padding
    sc = ops.SvgdScratch.allocate(n, "cuda")
    ops.svgd_pairdist_bandwidth(X, sc, 3e-4, 1.0, 50000.0)
    K, A = sc.K.cpu(), sc.A.cpu()
    cols = torch.randint(0, D, (4096,), generator=torch.Generator().manual_seed(2))
    cols = torch.cat([cols, torch.tensor([0, 1, 2, 3, D - 68, D - 65, D - 2, D - 1])]).cuda()
    x_old, g_cols = X[:, cols].cpu(), G[:, cols].cpu()
    ref_out = O.svgd_apply(x_old, g_cols, K, A).float()
    buf = torch.zeros(D, device="cuda")
    out_last = torch.empty(D, device="cuda")
    hyper = dict(lr=0.0, momentum=0.9, nesterov=True, weight_decay=3e-4)
    chk = X[:, ::4097].clone()
    ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=False, out_last=out_last, **hyper)
    assert torch.equal(X[:, ::4097], chk)  # lr = 0: particles bit-identical
    _, st = O.svgd_base_optimizer_steps(x_old, ref_out, "sgd", hyper)
    np.testing.assert_allclose(buf[cols].cpu().numpy(), st["momentum_buffer"].numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(out_last[cols].cpu().numpy(), ref_out[n - 1].numpy(), rtol=RTOL, atol=ATOL)
    hyper["lr"] = 0.05
    ops.svgd_apply_sgd(X, G, sc, buf, buf_initialized=True, out_last=out_last, **hyper)
    x_ref, st = O.svgd_base_optimizer_steps(x_old, ref_out, "sgd", hyper, st)
    check_update(X[:, cols].cpu().numpy(), x_old.numpy().astype(np.float64), x_ref.numpy(), "sgd full size")
    assert X[:, -64:].eq(0).all() and buf[-64:].eq(0).all()
    # Adam from step 0 on the updated particles
    del buf
    m, v = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.svgd_pairdist_bandwidth(X, sc, 3e-4, 1.0, 50000.0)
    x_old = X[:, cols].cpu()
    ref_out = O.svgd_apply(x_old, g_cols, sc.K.cpu(), sc.A.cpu()).float()
    ah = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    ops.svgd_apply_adam(X, G, sc, m, v, step0=0, lr=1e-3, out_last=out_last)
    x_ref, st = O.svgd_base_optimizer_steps(x_old, ref_out, "adam", ah)
    # the zero-gradient padding columns divide 0 by eps: they must stay exactly 0
    assert X[:, -64:].eq(0).all() and m[-64:].eq(0).all()
    live = (cols < D - 64).cpu()
    check_update(X[:, cols].cpu().numpy()[:, live], x_old.numpy().astype(np.float64)[:, live], x_ref.numpy()[:, live], "adam full size")


@pytest.fixture(params=[1, 2], ids=["direct", "tma"])
def ew_variant(request, cuda_lib):
    """Force the direct-LDG (1) or the TMA-staged (2) form of the elementwise kernels (ew_tma.cuh)."""
    assert cuda_lib.bde_tune(b"ew_variant", request.param) == 0
    yield request.param
    cuda_lib.bde_tune(b"ew_variant", 0)


# ---------------------------------------------------------------- SWAG
@pytest.mark.parametrize("D,K", [(501, 4), (4099, 10), (1 << 20, 10), (1237, 1), (1_000_003, 30), (3, 2), (2_500_001, 3)])
def test_swag_update_and_sample_vs_oracle(ops, ew_variant, D, K):
    g = torch.Generator().manual_seed(D + K)
    mean = torch.randn(D, generator=g) * 0.3
    sq = mean ** 2 + 0.01 * torch.rand(D, generator=g)
    ring = torch.zeros(K, D)
    d_mean, d_sq, d_ring = mean.cuda(), sq.cuda(), ring.cuda()
    updates = 0
    m_ref, s_ref = mean.clone(), sq.clone()
    for step in range(K + 3):  # wraps the ring
        theta = mean + 0.05 * torch.randn(D, generator=g)
        updates += 1
        m_ref, s_ref, col = O.swag_update(theta, m_ref, s_ref, updates)
        ring[(updates - 1) % K] = col
        ops.swag_update(theta.cuda(), d_mean, d_sq, d_ring[(updates - 1) % K], updates)
    # op-for-op identical arithmetic: bit-exact against the fp32 reference order
    assert torch.equal(d_mean.cpu(), m_ref) and torch.equal(d_sq.cpu(), s_ref) and torch.equal(d_ring.cpu(), ring)
    eps_k, eps_d = torch.randn(K, generator=g), torch.randn(D, generator=g)
    theta_out = torch.empty(D, device="cuda")
    ops.swag_sample(d_mean, d_sq, d_ring, updates % K, theta_out, eps_k=eps_k.cuda(), eps_d=eps_d.cuda())
    ref = O.swag_sample(m_ref, s_ref, O.swag_ring_to_reference(ring, updates), eps_k, eps_d)
    if K > 1:
        np.testing.assert_allclose(theta_out.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)
        ref64 = O.swag_sample(m_ref, s_ref, O.swag_ring_to_reference(ring, updates), eps_k, eps_d, dtype=torch.float64)
        np.testing.assert_allclose(theta_out.cpu().numpy(), ref64.numpy(), rtol=RTOL, atol=ATOL)
    else:  # K = 1 divides by sqrt(0) in the reference too: both are non-finite wherever dev != 0
        assert (~torch.isfinite(theta_out.cpu()) == ~torch.isfinite(ref)).all()


@pytest.mark.parametrize("S", [1, 2, 3, 5, 16, 17, 35])
@pytest.mark.parametrize("D,K,mis", [(100_003, 10, 0), (4099, 30, 0), (777, 3, 1), (1_000_000, 5, 0)])
def test_swag_sample_batch_equals_single_draws(ops, S, D, K, mis):
    """bde_swag_sample_batch: draw s == bde_swag_sample with stream_id + s, bit for bit, with Philox noise
    and with injected noise; ragged D, unaligned views, more draws than one launch holds (16), ring head != 0."""
    if S > 5 and D > 200_000:
        pytest.skip("large case covered at small S")
    g = torch.Generator().manual_seed(S * 7 + D)
    size = D + 8
    def vec():
        return torch.randn(size, generator=g).cuda()[mis:mis + D]
    mean = vec()
    sq = (mean ** 2 + 0.01 * torch.rand(D, generator=g).cuda()).contiguous() if mis == 0 else \
        (torch.zeros(size, device="cuda")[mis:mis + D].copy_(mean ** 2 + 0.01))
    ld = (D + 7) // 4 * 4 + (1 if mis else 0)
    dev = torch.as_strided(torch.randn(K * ld + 8, generator=g).cuda() * 0.1, (K, D), (ld, 1), storage_offset=mis)
    head = 2 % K
    ld_out = D + (3 if mis else 0)
    for injected in (False, True):
        ek = torch.randn(S, K, generator=g).cuda() if injected else None
        ed = torch.randn(S, D, generator=g).cuda() if injected else None
        out = torch.as_strided(torch.full((S * ld_out + 8,), float("nan"), device="cuda"), (S, D), (ld_out, 1), storage_offset=mis)
        ops.swag_sample_batch(mean, sq, dev, head, out, eps_k=ek, eps_d=ed, seed=99, stream_id=5)
        one = torch.zeros(size, device="cuda")[mis:mis + D]
        for s_ in range(S):
            ops.swag_sample(mean, sq, dev, head, one, eps_k=None if ek is None else ek[s_], eps_d=None if ed is None else ed[s_],
                            seed=99, stream_id=5 + s_)
            assert torch.equal(out[s_], one), (s_, injected)
        if S > 1:
            assert not torch.equal(out[0], out[1])


def test_swag_sample_matches_reference_fixture(ops, golden):
    g = golden("swag_steps.npz")
    D, K = g["deviations"].shape
    u = int(g["updates"])
    ring = torch.zeros(K, D)
    dev = torch.from_numpy(g["deviations"])
    for k in range(K):
        ring[(u % K + k) % K] = dev[:, k]
    off = 0
    for s in range(2):
        ek = torch.from_numpy(g["eps"][off:off + K]); off += K
        ed = torch.from_numpy(g["eps"][off:off + D]); off += D
        out = torch.empty(D, device="cuda")
        ops.swag_sample(torch.from_numpy(g["mean"]).cuda(), torch.from_numpy(g["sq"]).cuda(), ring.cuda(), u % K, out,
                        eps_k=ek.cuda(), eps_d=ed.cuda())
        np.testing.assert_allclose(out.cpu().numpy(), g["samples"][s], rtol=RTOL, atol=ATOL)


def test_swag_sample_philox_is_shard_independent(ops):
    D, K = 40_000, 6
    g = torch.Generator().manual_seed(3)
    mean, sq = torch.randn(D, generator=g).cuda(), (torch.rand(D, generator=g) + 1).cuda()
    ring = torch.randn(K, D, generator=g).cuda()
    full = torch.empty(D, device="cuda")
    ops.swag_sample(mean, sq, ring, 2, full, seed=77, stream_id=5)
    lo, hi = 12_032, 30_016
    part = torch.empty(hi - lo, device="cuda")
    ops.swag_sample(mean[lo:hi], sq[lo:hi], ring[:, lo:hi], 2, part, seed=77, stream_id=5, elem0=lo)
    assert torch.equal(part, full[lo:hi])  # same z_k on every rank, eps counted by global index
    # the in-kernel low-rank noise is the oracle's Philox stream
    zk = torch.from_numpy(O.philox_normal(K, 77, 5 ^ 0x5741))
    ed = torch.from_numpy(O.philox_normal(D, 77, 5))
    ref = O.swag_sample(mean.cpu(), sq.cpu(), O.swag_ring_to_reference(ring.cpu(), 2), zk, ed, dtype=torch.float64)
    np.testing.assert_allclose(full.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------- iVON
@pytest.mark.parametrize("D", [501, 4099, 1 << 20, 1_000_003, 2, 2_500_001])
def test_ivon_kernels_vs_oracle(ops, ew_variant, D):
    g = torch.Generator().manual_seed(D)
    mean = 0.3 * torch.randn(D, generator=g)
    prec = 10.0 / 768 + 0.01 * torch.rand(D, generator=g)
    prec[:min(3, D)] = torch.tensor([1e-6, 1e-4, 0.0])[:D]  # below / at the clamp
    mom = 0.01 * torch.randn(D, generator=g)
    N, S = 768.0, 3
    d_mean, d_prec, d_mom = mean.cuda(), prec.cuda(), mom.cuda()
    d_dsum, d_theta, d_acc = (torch.full((D,), float("nan"), device="cuda") for _ in range(3))
    dsum_ref, acc_ref = None, None
    for s in range(S):
        eps = torch.randn(D, generator=g)
        ops.ivon_sample(d_mean, d_prec, d_dsum, d_theta, n_eff=N, first=(s == 0), eps=eps.cuda())
        th_ref, dsum_ref = O.ivon_sample(mean, prec, dsum_ref, eps, N)
        np.testing.assert_allclose(d_theta.cpu().numpy(), th_ref.numpy(), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(d_dsum.cpu().numpy(), dsum_ref.numpy(), rtol=RTOL, atol=ATOL)
        dsum_ref = d_dsum.cpu().clone()  # continue from the device state so the update check below stays tight
        grad = 1e-2 * torch.randn(D, generator=g)
        ops.ivon_accumulate(d_acc, grad.cuda(), first=(s == 0))
        acc_ref = grad if acc_ref is None else acc_ref + grad
        assert torch.equal(d_acc.cpu(), acc_ref)
    kw = dict(mc_samples=S, step=4, lr=1e-2, prior_prec=10.0, n_eff=N, tempering=0.7, damping=1e-3)
    ops.ivon_update(d_acc, d_dsum, d_mean, d_mom, d_prec, beta1=0.9, beta2=0.999, **kw)
    m_ref, mo_ref, p_ref = O.ivon_update(acc_ref, dsum_ref, mean, mom, prec, betas=(0.9, 0.999), **kw)
    ok = torch.isfinite(p_ref)  # prec == 0 element divides by zero in the reference as well
    for got, ref in ((d_mean, m_ref), (d_mom, mo_ref), (d_prec, p_ref)):
        np.testing.assert_allclose(got.cpu()[ok].numpy(), ref[ok].numpy(), rtol=RTOL, atol=ATOL)
    # deterministic mode: theta = mean exactly, delta = 0
    ops.ivon_sample(d_mean, d_prec, d_dsum, d_theta, n_eff=N, first=True, deterministic=True)
    assert torch.equal(d_theta, d_mean) and d_dsum.eq(0).all()


@pytest.mark.parametrize("S,first,det", [(1, True, False), (3, True, False), (4, False, False), (7, False, False), (3, True, True),
                                         (16, False, False), (19, True, False), (31, False, False)])
@pytest.mark.parametrize("D,mis", [(100_003, 0), (4099, 0), (777, 1), (3_000_000, 0)])
def test_ivon_sample_batch_equals_single_draws(ops, S, first, det, D, mis):
    """bde_ivon_sample_batch: draw s == bde_ivon_sample with stream_id + s * stride, and delta_sum ends as after S
    single calls, bit for bit — Philox and injected noise, first / continuing accumulation, deterministic groups,
    ragged D, unaligned views, D large enough for the TMA-staged single-draw kernel.  Aligned Philox cases run the fast
    kernel in passes of 16 / 8 / 4 / 2 draws (+ the general kernel for an odd last draw and the D % 4 tail)."""
    g = torch.Generator().manual_seed(S * 13 + D)
    def vec(scale=1.0, shift=0.0):
        return (torch.randn(D + 8, generator=g) * scale + shift).cuda()[mis:mis + D]
    mean, prec = vec(0.1), vec(1e-3, 0.02).abs_()
    ld_out = D + (3 if mis else 0)
    for injected in (False, True):
        eps = torch.randn(S, D, generator=g).cuda() if injected else None
        ds0 = vec(0.3)
        ds_b, ds_s = ds0.clone(), ds0.clone()
        out = torch.as_strided(torch.full((S * ld_out + 8,), float("nan"), device="cuda"), (S, D), (ld_out, 1), storage_offset=mis)
        ops.ivon_sample_batch(mean, prec, ds_b, out, n_eff=768.0, first=first, deterministic=det, eps=eps, seed=7,
                              stream_id=11, stream_stride=2)
        one = torch.zeros(D + 8, device="cuda")[mis:mis + D]
        for s_ in range(S):
            ops.ivon_sample(mean, prec, ds_s, one, n_eff=768.0, first=first and s_ == 0, deterministic=det,
                            eps=None if eps is None else eps[s_], seed=7, stream_id=11 + 2 * s_)
            assert torch.equal(out[s_], one), (s_, injected)
        assert torch.equal(ds_b, ds_s)
        if S > 1 and not det:
            assert not torch.equal(out[0], out[1])


def test_ivon_sample_philox_matches_oracle_stream(ops, ew_variant):
    D = 100_003
    mean = torch.zeros(D, device="cuda")
    prec = torch.full((D,), 1.0 / 768, device="cuda")
    dsum, theta = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    ops.ivon_sample(mean, prec, dsum, theta, n_eff=768.0, first=True, seed=1234, stream_id=9)
    z = O.philox_normal(D, 1234, 9)
    np.testing.assert_allclose(theta.cpu().numpy(), z, rtol=1e-4, atol=1e-4)
    lo = 50_048
    part = torch.empty(D - lo, device="cuda")
    ops.ivon_sample(mean[lo:], prec[lo:], dsum[lo:].clone(), part, n_eff=768.0, first=True, seed=1234, stream_id=9, elem0=lo)
    assert torch.equal(part, theta[lo:])


def test_philox_normal_kernel(ops):
    n = 1 << 22
    out = torch.empty(n, device="cuda")
    ops.philox_normal(out, seed=42, stream_id=3)
    z = out.cpu().numpy()
    np.testing.assert_allclose(z[:200_000], O.philox_normal(200_000, 42, 3), rtol=1e-4, atol=1e-4)  # MUFU log2/sin/cos
    assert abs(z.mean()) < 2e-3 and abs(z.std() - 1) < 2e-3
    assert abs(((z ** 3).mean())) < 1e-2 and abs((z ** 4).mean() - 3) < 3e-2
    out2 = torch.empty(n, device="cuda")
    ops.philox_normal(out2, seed=42, stream_id=4)
    assert abs(np.corrcoef(z[:100000], out2.cpu().numpy()[:100000])[0, 1]) < 0.02


# ---------------------------------------------------------------- BBB
def test_gauss_and_kl_kernels_vs_reference_vectors(ops, golden):
    g = golden("vectors.npz")
    c = lambda k: torch.from_numpy(g[k]).cuda()
    mu, rho, eps = c("mu"), c("rho"), c("eps")
    w = torch.empty_like(mu)
    ops.gauss_sample_fwd(mu, rho, w, eps=eps)
    np.testing.assert_allclose(w.cpu().numpy(), g["w"], rtol=RTOL, atol=ATOL)
    grho = torch.empty_like(mu)
    ops.gauss_sample_bwd(c("grad_w"), rho, grho, eps=eps)
    np.testing.assert_allclose(grho.cpu().numpy(), g["grad_rho"], rtol=RTOL, atol=ATOL)
    ws = ops.value_workspace("cuda")
    val = torch.zeros((), dtype=torch.float64, device="cuda")
    gmu, gr = torch.empty_like(mu), torch.empty_like(mu)
    ops.kl_gauss(mu, rho, 0.5, 0.8, value=val, grad_mu=gmu, grad_rho=gr, ws=ws)
    np.testing.assert_allclose(val.item(), g["kl_gauss"], rtol=RTOL)
    np.testing.assert_allclose(gmu.cpu().numpy(), g["kl_gauss_gmu"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(gr.cpu().numpy(), g["kl_gauss_grho"], rtol=RTOL, atol=ATOL)
    # accumulate + device-side scale: g += 2 * 0.25 * dKL
    scale = torch.tensor([0.25], device="cuda")
    ops.kl_gauss(mu, rho, 0.5, 0.8, grad_mu=gmu, grad_rho=gr, grad_scale=2.0, grad_scale_dev=scale, accumulate=True)
    np.testing.assert_allclose(gmu.cpu().numpy(), 1.5 * g["kl_gauss_gmu"], rtol=RTOL, atol=2 * ATOL)
    mmu = c("mix_mu")
    mval = torch.zeros((), dtype=torch.float64, device="cuda")
    mg = torch.empty_like(mmu)
    ops.kl_mixture(mmu, 0.3, 1.0, 0.0025, value=mval, grad_mu=mg, ws=ws)
    np.testing.assert_allclose(mval.item(), g["kl_mix"], rtol=RTOL)
    np.testing.assert_allclose(mg.cpu().numpy(), g["kl_mix_gmu"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("P", [1, 7, 592_130, 1_000_003])
def test_kl_and_l2_values_vs_oracle_sizes(ops, P):
    g = torch.Generator().manual_seed(P)
    mu, rho = 0.1 * torch.randn(P, generator=g), -3 + 0.5 * torch.randn(P, generator=g)
    ws = ops.value_workspace("cuda")
    val = torch.zeros((), dtype=torch.float64, device="cuda")
    gmu, gr = torch.empty(P, device="cuda"), torch.empty(P, device="cuda")
    ops.kl_gauss(mu.cuda(), rho.cuda(), 0.0, 1.0, value=val, grad_mu=gmu, grad_rho=gr, ws=ws)
    v_ref, gm_ref, gr_ref = O.kl_gauss(mu, rho, 0.0, 1.0)
    np.testing.assert_allclose(val.item(), v_ref.item(), rtol=RTOL)
    np.testing.assert_allclose(gmu.cpu().numpy(), gm_ref.numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(gr.cpu().numpy(), gr_ref.numpy(), rtol=RTOL, atol=ATOL * 30)  # |d/drho| ~ 20
    # value-only launch gives the same value and is deterministic
    val2 = torch.zeros((), dtype=torch.float64, device="cuda")
    ops.kl_gauss(mu.cuda(), rho.cuda(), 0.0, 1.0, value=val2, ws=ws)
    assert val2.item() == val.item()
    l2v = torch.zeros((), dtype=torch.float64, device="cuda")
    grad = torch.ones(P, device="cuda")
    ops.l2_term(mu.cuda(), 0.01, value=l2v, grad=grad, grad_scale=3.0, accumulate=True, ws=ws)
    lv_ref, lg_ref = O.l2_term(mu, 0.01)
    np.testing.assert_allclose(l2v.item(), lv_ref.item(), rtol=RTOL)
    np.testing.assert_allclose(grad.cpu().numpy(), 1.0 + 3.0 * lg_ref.numpy(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("prior_kind", ["gauss", "mixture"])
@pytest.mark.parametrize("accumulate", [False, True])
def test_prior_terms_multi_tensor_vs_oracle(ops, prior_kind, accumulate):
    """bde_prior_terms_value_and_grad over a list of 130 tensors (three table chunks): Gaussian parameters of
    ragged sizes, 4-byte-misaligned views, deterministic tensors with their own l2 scales, tensors without a
    gradient target — value and gradients against the oracle, and bit for bit against the per-tensor kernels."""
    g = torch.Generator().manual_seed(7)
    sizes = [1, 3, 4, 5, 768 * 2, 592_130 // 8, 17, 64, 1001] + [int(x) for x in torch.randint(1, 3000, (121,), generator=g)]
    kinds, a, b, l2, ga, gb = [], [], [], [], [], []
    code = ops.PRIOR_GAUSS if prior_kind == "gauss" else ops.PRIOR_MIXTURE
    prior = (0.3, 0.8, 0.0) if prior_kind == "gauss" else (0.4, 1.5, 0.1)

    def dev(t, mis):
        buf = torch.zeros(t.numel() + 4, device="cuda")
        v = buf[mis:mis + t.numel()]
        v.copy_(t)
        return v

    for i, P in enumerate(sizes):
        mis = 1 if i % 7 == 3 else 0
        if i % 3 == 2:
            kinds.append(ops.PRIOR_L2)
            a.append(dev(0.05 * torch.randn(P, generator=g), mis))
            b.append(None)
            l2.append(0.01 * (1 + i % 4))
        else:
            kinds.append(code)
            a.append(dev(0.3 * torch.randn(P, generator=g), mis))
            b.append(dev(-3 + 0.5 * torch.randn(P, generator=g), mis) if code == ops.PRIOR_GAUSS else None)
            l2.append(0.0)
        want = i % 11 != 5   # some tensors get no gradient
        init = 0.1 * torch.randn(P, generator=g)
        ga.append(dev(init, mis) if want else None)
        gb.append(dev(init * 2, mis) if (want and kinds[-1] == ops.PRIOR_GAUSS) else None)
    ga0 = [None if t is None else t.clone() for t in ga]
    gb0 = [None if t is None else t.clone() for t in gb]
    value = torch.zeros((), dtype=torch.float64, device="cuda")
    scale_dev = torch.tensor([0.25], device="cuda")
    ws = ops.value_workspace("cuda")
    ops.prior_terms(kinds, a, b, l2_scales=l2, prior=prior, value=value, grad_a=ga, grad_b=gb, grad_scale=2.0,
                    grad_scale_dev=scale_dev, accumulate=accumulate, ws=ws)
    total = 0.0
    v1 = torch.zeros((), dtype=torch.float64, device="cuda")
    for i, P in enumerate(sizes):
        x = a[i].cpu()
        if kinds[i] == ops.PRIOR_GAUSS:
            val, g1, g2 = O.kl_gauss(x, b[i].cpu(), prior[0], prior[1])
        elif kinds[i] == ops.PRIOR_MIXTURE:
            val, g1 = O.kl_mixture(x, *prior)
            g2 = None
        else:
            val, g1 = O.l2_term(x, l2[i])
            g2 = None
        total += float(val)
        if ga[i] is None:
            continue
        base1 = ga0[i].cpu() if accumulate else 0.0
        np.testing.assert_allclose(ga[i].cpu().numpy(), (base1 + 0.5 * g1).numpy(), rtol=RTOL, atol=ATOL, err_msg=f"tensor {i}")
        if g2 is not None:
            base2 = gb0[i].cpu() if accumulate else 0.0
            np.testing.assert_allclose(gb[i].cpu().numpy(), (base2 + 0.5 * g2).numpy(), rtol=RTOL, atol=30 * ATOL, err_msg=f"tensor {i}")
        # the per-tensor kernels do the same arithmetic
        r1 = ga0[i].clone()
        if kinds[i] == ops.PRIOR_GAUSS:
            r2 = gb0[i].clone()
            ops.kl_gauss(a[i], b[i], prior[0], prior[1], value=v1, grad_mu=r1, grad_rho=r2, grad_scale=2.0,
                         grad_scale_dev=scale_dev, accumulate=accumulate, ws=ws)
            assert torch.equal(r2, gb[i])
        elif kinds[i] == ops.PRIOR_MIXTURE:
            ops.kl_mixture(a[i], *prior, value=v1, grad_mu=r1, grad_scale=2.0, grad_scale_dev=scale_dev,
                           accumulate=accumulate, ws=ws)
        else:
            ops.l2_term(a[i], l2[i], value=v1, grad=r1, grad_scale=2.0, grad_scale_dev=scale_dev, accumulate=accumulate, ws=ws)
        assert torch.equal(r1, ga[i]), f"tensor {i}"
    np.testing.assert_allclose(value.item(), total, rtol=1e-6)
    # value only, and an empty list
    v2 = torch.ones((), dtype=torch.float64, device="cuda")
    ops.prior_terms(kinds, a, b, l2_scales=l2, prior=prior, value=v2, ws=ws)
    assert v2.item() == value.item()   # deterministic reduction
    ops.prior_terms([], [], [], value=v2, ws=ws)
    assert v2.item() == 0.0


def test_gauss_sample_philox_fwd_bwd_consistent(ops):
    P = 100_001
    mu, rho = torch.zeros(P, device="cuda"), torch.full((P,), 0.5413, device="cuda")  # softplus = 1.0000
    w = torch.empty(P, device="cuda")
    ops.gauss_sample_fwd(mu, rho, w, seed=5, stream_id=11)
    z = O.philox_normal(P, 5, 11)
    sig = torch.nn.functional.softplus(rho[0]).item()
    np.testing.assert_allclose(w.cpu().numpy() / sig, z, rtol=1e-4, atol=1e-4)
    grho = torch.empty(P, device="cuda")
    ops.gauss_sample_bwd(torch.ones(P, device="cuda"), rho, grho, seed=5, stream_id=11)
    np.testing.assert_allclose(grho.cpu().numpy(), (w / sig * torch.sigmoid(rho)).cpu().numpy(), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- gather / scatter
def test_multi_tensor_copy_modes(ops):
    from beyond_deep_ensembles_b200.layout import ParamLayout
    g = torch.Generator().manual_seed(0)
    shapes = [(50, 8), (50,), (1, 50), (1,), (3, 3, 3, 3), (129,)] * 40  # 240 tensors -> 3 table chunks
    tensors = [torch.randn(s, generator=g).cuda() for s in shapes]
    L = ParamLayout(tensors)
    row = torch.full((L.size,), 7.0, device="cuda")
    ops.multi_tensor_copy(row, tensors, L.offsets, mode=0)
    for v, t in zip(L.views(row), tensors):
        assert torch.equal(v, t)
    ops.multi_tensor_copy(row, tensors, L.offsets, mode=1)
    for v, t in zip(L.views(row), tensors):
        assert torch.equal(v, t + t)
    outs = [torch.zeros_like(t) for t in tensors]
    ops.multi_tensor_copy(row, outs, L.offsets, mode=2)
    for o, t in zip(outs, tensors):
        assert torch.equal(o, t + t)
    # unpadded (logical) offsets exercise the unaligned scalar path
    offs = np.cumsum([0] + [t.numel() for t in tensors[:-1]]).tolist()
    flat = torch.zeros(sum(t.numel() for t in tensors), device="cuda")
    ops.multi_tensor_copy(flat, tensors, offs, mode=0)
    assert torch.equal(flat, torch.cat([t.reshape(-1) for t in tensors]))


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("scale,bad", [(1024.0, None), (1.0, None), (65536.0, "inf"), (8.0, "nan")])
def test_multi_tensor_unscale_copy_vs_torch(ops, mode, scale, bad):
    """bde_multi_tensor_unscale_copy == GradScaler's _amp_foreach_non_finite_check_and_unscale_ followed by the
    gather / gather-add, bit for bit; found_inf raised exactly when a source value is not finite; sources untouched;
    ragged sizes, an unaligned tensor, more tensors than one kernel-parameter table holds (112)."""
    g = torch.Generator().manual_seed(17)
    sizes = [1, 3, 4, 5, 64, 1000, 4097] + [7] * 120
    tensors = [torch.randn(sz, generator=g).cuda() for sz in sizes]
    tensors[5] = torch.randn(1001, generator=g).cuda()[1:]          # 4-byte-aligned only
    if bad:
        tensors[3][2] = float(bad)
    before = [t.clone() for t in tensors]
    offsets, off = [], 0
    for sz in sizes:
        offsets.append(off)
        off += -(-sz // 64) * 64
    flat0 = torch.randn(off, generator=g).cuda()
    inv = torch.tensor(1.0 / scale, device="cuda")
    found = torch.zeros((), device="cuda")
    flat = flat0.clone()
    ops.multi_tensor_copy(flat, tensors, offsets, mode, inv_scale=inv, found_inf=found)
    # torch's own kernel on copies, then the plain gather
    ref_t = [t.clone() for t in tensors]
    ref_found = torch.zeros(1, device="cuda")
    torch._amp_foreach_non_finite_check_and_unscale_(ref_t, ref_found, inv)
    ref = flat0.clone()
    ops.multi_tensor_copy(ref, ref_t, offsets, mode)
    assert found.item() == ref_found.item() == (1.0 if bad else 0.0)
    assert torch.equal(torch.nan_to_num(flat, nan=123.0), torch.nan_to_num(ref, nan=123.0))
    for t, b in zip(tensors, before):
        assert torch.equal(torch.nan_to_num(t, nan=5.0), torch.nan_to_num(b, nan=5.0))


def test_host_buffer_step_matches_device_step(ops):
    """The end-to-end host-buffer path (chunked H2D -> K1 .. K1b -> K2 -> D2H pipeline) equals the
    oracle, for a ragged D, a chunk size that does not divide it, and pageable as well as pinned memory."""
    n, D, chunk = 10, 1_000_003, 65_536
    X, G = particles(n, D, 21)
    ref, info = O.svgd_step_fused(X, G, 0.01, 1.0, 50000.0)
    for pin in (True, False):
        Xh, Gh = (X.pin_memory(), G.pin_memory()) if pin else (X, G)
        outh = torch.empty_like(X).pin_memory() if pin else torch.empty_like(X)
        st = ops.HostStaging.allocate(n, D, chunk, "cuda")
        sc = ops.SvgdScratch.allocate(n, "cuda")
        torch.cuda.synchronize()
        ops.svgd_step_host(Xh, Gh, outh, st, sc, 0.01, 1.0, 50000.0)
        assert tuple(sc.sel.cpu().tolist()) == info["sel"]
        np.testing.assert_allclose(sc.info[0].item(), info["h"], rtol=1e-6)
        np.testing.assert_allclose(outh.numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)


# ---------------------------------------------------------------- K1 for n = 16 / 20: centred-Gram kernel + exact redo
def _gram_case(ops, cuda_lib, X, variant=3, guard_x1000=0):
    n, D = X.shape
    dX = dev_matrix(X, (D + 3) // 4 * 4)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    cuda_lib.bde_tune(b"pairdist_variant", variant)
    cuda_lib.bde_tune(b"gram_guard_x1000", guard_x1000)
    try:
        ops.svgd_pairdist_bandwidth(dX, sc, 0.01, 1.0, 50000.0)
        first = (sc.dist.clone(), sc.K.clone(), sc.A.clone(), sc.sel.clone())
        ops.svgd_pairdist_bandwidth(dX, sc, 0.01, 1.0, 50000.0)   # ring / ticket / flag clean across launches
        assert torch.equal(sc.dist, first[0]) and torch.equal(sc.K, first[1]) and torch.equal(sc.A, first[2])
    finally:
        cuda_lib.bde_tune(b"pairdist_variant", 0)
        cuda_lib.bde_tune(b"gram_guard_x1000", 0)
    return sc


def _check_vs_fp64(sc, X):
    d_ref = O.svgd_pairdist(X)
    bw = O.svgd_bandwidth(d_ref, 0.01, 1.0, 50000.0)
    np.testing.assert_allclose(sc.dist.cpu().numpy(), d_ref.numpy(), rtol=2e-6, atol=1e-12)
    assert tuple(sc.sel.cpu().tolist()) == bw["sel"]
    np.testing.assert_allclose(sc.K.cpu().numpy(), bw["K"].numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(sc.A.cpu().numpy(), bw["A"].numpy(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n,D", [(20, 2_000_003), (20, 1_048_576), (20, 300_001), (16, 2_000_003), (16, 777_777), (20, 255),
                                 (16, 4)])
def test_pairgram_vs_oracle(ops, cuda_lib, n, D):
    """Centred-Gram K1 (svgd_gram.cuh) against the fp64 oracle: distances to 2e-6, identical median selection, K / A
    within the north-star tolerance; ragged D (TMA zero fill) and a D smaller than one tile included."""
    X, _ = particles(n, D, seed=5 * n + D)
    sc = _gram_case(ops, cuda_lib, X)
    assert not sc.exact_redo()
    _check_vs_fp64(sc, X)


@pytest.mark.parametrize("n", [16, 20])
def test_pairgram_common_offset_is_harmless(ops, cuda_lib, n):
    """Particles that share a large common offset (a trained network's weights) and differ by small perturbations:
    the uncentred Gram form would lose every digit; centring on particle 0 keeps the result inside 2e-6."""
    D = 1_500_000
    g = torch.Generator().manual_seed(n)
    base = torch.randn(1, D, generator=g)
    X = base + 1e-3 * (1 + 0.1 * torch.arange(n, dtype=torch.float32)).unsqueeze(1) * torch.randn(n, D, generator=g)
    sc = _gram_case(ops, cuda_lib, X)
    assert not sc.exact_redo()
    _check_vs_fp64(sc, X)


@pytest.mark.parametrize("n", [16, 20])
@pytest.mark.parametrize("case", ["duplicate", "near-duplicate", "outlier-0", "forced"])
def test_pairgram_guard_falls_back_to_exact_distances(ops, cuda_lib, n, case):
    """When some pair is much closer to each other than to particle 0 the guard raises `redo` and the direct kernel
    queued behind recomputes: results are then bit-identical to the direct variant (duplicates give exact zeros)."""
    D = 1_200_003
    X, _ = particles(n, D, seed=11 * n)
    guard = 0
    if case == "duplicate":
        X[7] = X[5]
    elif case == "near-duplicate":
        X[7] = X[5] + 1e-4 * X[3]
    elif case == "outlier-0":
        X[0] = 200.0 * X[0]
    else:
        guard = 1
    sc = _gram_case(ops, cuda_lib, X, guard_x1000=guard)
    assert sc.exact_redo()
    direct = _gram_case(ops, cuda_lib, X, variant=2)
    assert torch.equal(sc.dist, direct.dist) and torch.equal(sc.K, direct.K) and torch.equal(sc.A, direct.A)
    assert torch.equal(sc.sel, direct.sel)
    if case == "duplicate":
        assert float(sc.dist[5, 7]) == 0.0
    _check_vs_fp64(sc, X)


def test_pairgram_is_the_default_at_large_D(ops):
    """Auto selection: n = 20 at a D with >= 4 tiles per SM runs the Gram kernel (redo flag cleared by it) and agrees
    with the fp64 evaluation on the device; the n = 20 full-size step stays inside the tolerance."""
    n, D = 20, 20_000_000
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(n, D, device="cuda", generator=g)
    X *= (0.05 * (1 + 0.1 * torch.arange(n, device="cuda", dtype=torch.float32))).unsqueeze(1)
    sc = ops.SvgdScratch.allocate(n, "cuda")
    sc.ws[1] = (1 << 32) | int(sc.ws[1].item())   # pre-set redo: the Gram kernel must clear it
    ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)
    assert not sc.exact_redo()
    ref = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    step = 1 << 21
    for c0 in range(0, D, step):
        x = X[:, c0:c0 + step].double()
        ref += torch.cdist(x, x, p=2, compute_mode="donot_use_mm_for_euclid_dist") ** 2
    np.testing.assert_allclose(sc.dist.cpu().numpy(), ref.cpu().numpy(), rtol=2e-6)
    bw = O.svgd_bandwidth(ref.cpu(), 0.01, 1.0, 50000.0)
    assert tuple(sc.sel.cpu().tolist()) == bw["sel"]


# ---------------------------------------------------------------- BASELINE sizes of the elementwise family
def _sample_cols(D, count=4096, seed=1):
    cols = torch.randint(0, D, (count,), generator=torch.Generator().manual_seed(seed))
    return torch.unique(torch.cat([cols, torch.arange(64), torch.arange(D - 64, D)]))


def test_swag_full_size_properties(ops):
    """C3 (iWildCam ResNet-50 + fc182: D = 23,880,950, K = 10): K + 2 updates (the ring wraps) are bit-exact against the
    fp32 oracle on a column sample (4096 random columns, the first and the last 64), a draw matches the oracle on the
    same columns, the single and the batched sampler agree bit for bit on the WHOLE vector, and an update with
    theta = mean leaves mean unchanged and writes an all-zero deviation row (size-independent properties)."""
    D, K = 23_880_950, 10
    g = torch.Generator(device="cuda").manual_seed(3)
    mean = torch.randn(D, device="cuda", generator=g) * 0.3
    sq = mean * mean + 0.01 * torch.rand(D, device="cuda", generator=g)
    ring = torch.zeros(K, D, device="cuda")
    cols = _sample_cols(D)
    cd = cols.cuda()
    m_ref, s_ref = mean[cd].cpu(), sq[cd].cpu()
    ring_ref = torch.zeros(K, cols.numel())
    updates = 0
    for _ in range(K + 2):
        theta = mean + 0.05 * torch.randn(D, device="cuda", generator=g)
        updates += 1
        th_c = theta[cd].cpu()
        ops.swag_update(theta, mean, sq, ring[(updates - 1) % K], updates)
        m_ref, s_ref, col = O.swag_update(th_c, m_ref, s_ref, updates)
        ring_ref[(updates - 1) % K] = col
    assert torch.equal(mean[cd].cpu(), m_ref) and torch.equal(sq[cd].cpu(), s_ref) and torch.equal(ring[:, cd].cpu(), ring_ref)
    # one draw with injected noise on the sampled columns
    eps_k = torch.randn(K, generator=torch.Generator().manual_seed(5))
    eps_d = torch.randn(D, device="cuda", generator=g)
    out = torch.empty(D, device="cuda")
    ops.swag_sample(mean, sq, ring, updates % K, out, eps_k=eps_k.cuda(), eps_d=eps_d)
    ref = O.swag_sample(m_ref, s_ref, O.swag_ring_to_reference(ring_ref, updates), eps_k, eps_d[cd].cpu(), dtype=torch.float64)
    np.testing.assert_allclose(out[cd].cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)
    # Philox draws: batched == single, whole vector, bit for bit
    outs = torch.empty(3, D, device="cuda")
    ops.swag_sample_batch(mean, sq, ring, updates % K, outs, seed=11, stream_id=40)
    for s_ in range(3):
        ops.swag_sample(mean, sq, ring, updates % K, out, seed=11, stream_id=40 + s_)
        assert torch.equal(outs[s_], out)
    # idempotence-like property: theta == mean  ->  mean unchanged, deviation row exactly zero
    before = mean.clone()
    updates += 1
    ops.swag_update(before.clone(), mean, sq, ring[(updates - 1) % K], updates)
    torch.testing.assert_close(mean, before, rtol=3e-7, atol=0)     # (u m + m) / (u + 1) rounds at most twice
    assert float(ring[(updates - 1) % K].abs().max()) <= 3e-7 * float(before.abs().max())


def test_ivon_full_size_properties(ops):
    """C4b (CivilComments DistilBERT + head, full-model iVON: D = 66,955,010): two MC samples + gradient accumulation +
    the update against the oracle on a column sample; the accumulate kernel is exact; batched == single draws on the
    whole vector; a deterministic draw returns the mean exactly."""
    D = 66_955_010
    N, S = 269038.0, 2
    g = torch.Generator(device="cuda").manual_seed(4)
    mean = torch.randn(D, device="cuda", generator=g) * 0.05
    prec = torch.rand(D, device="cuda", generator=g) * 1e-4 + 10.0 / 269038
    mom = torch.randn(D, device="cuda", generator=g) * 1e-4
    dsum, theta, acc = (torch.empty(D, device="cuda") for _ in range(3))
    cols = _sample_cols(D)
    cd = cols.cuda()
    mean_c, prec_c, mom_c = mean[cd].cpu(), prec[cd].cpu(), mom[cd].cpu()
    dsum_ref, acc_ref = None, None
    for s in range(S):
        eps = torch.randn(D, device="cuda", generator=g)
        ops.ivon_sample(mean, prec, dsum, theta, n_eff=N, first=(s == 0), eps=eps)
        th_ref, dsum_ref = O.ivon_sample(mean_c, prec_c, dsum_ref, eps[cd].cpu(), N)
        np.testing.assert_allclose(theta[cd].cpu().numpy(), th_ref.numpy(), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(dsum[cd].cpu().numpy(), dsum_ref.numpy(), rtol=RTOL, atol=ATOL)
        dsum_ref = dsum[cd].cpu().clone()
        grad = torch.randn(D, device="cuda", generator=g) * 1e-3
        ops.ivon_accumulate(acc, grad, first=(s == 0))
        acc_ref = grad[cd].cpu() if acc_ref is None else acc_ref + grad[cd].cpu()
        assert torch.equal(acc[cd].cpu(), acc_ref)
    kw = dict(mc_samples=S, step=7, lr=1e-5, prior_prec=10.0, n_eff=N, tempering=1.0, damping=1e-3)
    ops.ivon_update(acc, dsum, mean, mom, prec, beta1=0.9, beta2=0.999, **kw)
    m_ref, mo_ref, p_ref = O.ivon_update(acc_ref, dsum_ref, mean_c, mom_c, prec_c, betas=(0.9, 0.999), **kw)
    for got, ref in ((mean, m_ref), (mom, mo_ref), (prec, p_ref)):
        np.testing.assert_allclose(got[cd].cpu().numpy(), ref.numpy(), rtol=RTOL, atol=ATOL)
    assert bool(torch.isfinite(prec).all()) and bool(torch.isfinite(mean).all())
    # batched == single (Philox), whole vector
    outs = torch.empty(2, D, device="cuda")
    ds_b, ds_s = dsum.clone(), dsum.clone()
    ops.ivon_sample_batch(mean, prec, ds_b, outs, n_eff=N, first=False, seed=5, stream_id=21, stream_stride=3)
    for s_ in range(2):
        ops.ivon_sample(mean, prec, ds_s, theta, n_eff=N, first=False, seed=5, stream_id=21 + 3 * s_)
        assert torch.equal(outs[s_], theta)
    assert torch.equal(ds_b, ds_s)
    ops.ivon_sample(mean, prec, dsum, theta, n_eff=N, first=True, deterministic=True)
    assert torch.equal(theta, mean) and bool(dsum.eq(0).all())


@pytest.mark.parametrize("kind,n,D", [("adam", 10, 40_003), ("adamw", 5, 300_000), ("adam", 20, 20_000)])
def test_fused_adam_does_not_drift_over_200_steps(ops, kind, n, D):
    """the fused Adam uses MUFU sqrt / rcp (~1 ulp each).  200 consecutive SVGD steps (200 n Adam steps on
    the shared state) next to `torch.optim.Adam` stepped literally as the reference does (svgd.py:92-103), both fed
    the SAME new gradients every step: the particles stay within the north-star tolerance relative to the size of
    the accumulated movement, i.e. the approximate instructions do not accumulate a bias."""
    g = torch.Generator().manual_seed(n + D)
    X0 = (0.05 * torch.randn(n, D, generator=g)).cuda()
    hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01 if kind == "adamw" else 0.0)
    Xf = X0.clone()
    sc = ops.SvgdScratch.allocate(n, "cuda")
    m = torch.zeros(D, device="cuda")
    v = torch.zeros(D, device="cuda")
    # literal torch side: ONE parameter re-pointed at particle i, one optimizer whose state is shared
    Xt = X0.clone()
    param = torch.nn.Parameter(Xt[0])
    cls = torch.optim.AdamW if kind == "adamw" else torch.optim.Adam
    base = cls([param], foreach=False, **hyper)
    out = torch.empty_like(X0)
    worst = 0.0
    for step in range(200):
        G = (1e-2 * torch.randn(n, D, generator=g)).cuda()
        # the kernel matrix of the FUSED trajectory drives both sides, so the comparison isolates the optimizer arithmetic
        ops.svgd_pairdist_bandwidth(Xf, sc, 0.01, 1.0, 50000.0)
        ops.svgd_apply(Xf, G, out, sc)
        ops.svgd_apply_adam(Xf, G, sc, m, v, step0=step * n, lr=hyper["lr"], beta1=0.9, beta2=0.999, eps=1e-8,
                            weight_decay=hyper["weight_decay"], decoupled_weight_decay=(kind == "adamw"))
        for i in range(n):
            param.data = Xt[i]
            param.grad = out[i].clone()
            base.step()
        if step % 20 == 19 or step == 0:
            moved = (Xt - X0).abs().max().item()
            diff = (Xf - Xt).abs().max().item()
            worst = max(worst, diff / max(moved, 1e-30))
            # per-step-relative bound: the gap never exceeds 1e-5 of the distance travelled (+ 1e-6 absolute)
            assert diff <= 1e-5 * moved + 1e-6, (step, diff, moved)
    np.testing.assert_allclose(Xf.cpu().numpy(), Xt.cpu().numpy(), rtol=RTOL, atol=ATOL)
    st = base.state[param]
    np.testing.assert_allclose(m.cpu().numpy(), st["exp_avg"].cpu().numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(v.cpu().numpy(), st["exp_avg_sq"].cpu().numpy(), rtol=1e-4, atol=1e-9)
