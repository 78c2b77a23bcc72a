"""Pin the oracle (oracle/bde_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py -> tests/golden/*.npz).  CPU only."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import ATOL, RTOL
from oracle import bde_oracle as O

RBF_CASES = [(5, 37), (10, 501), (20, 1000), (3, 64), (2, 9), (10, 4099), (16, 257), (1, 33)]


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    return x if dtype is None else x.to(dtype)


@pytest.mark.parametrize("n,D", RBF_CASES)
def test_pairdist_and_median_match_reference(golden, n, D):
    g = golden("rbf.npz")
    key = f"n{n}_D{D}"
    X = t(g[f"{key}_X"])
    d = O.svgd_pairdist(X)
    # reference cdist**2 in fp64 (sqrt then square): agree to fp64 rounding
    np.testing.assert_allclose(torch.sqrt(d).pow(2).numpy(), g[f"{key}_d64"], rtol=1e-12, atol=1e-15)
    bw = O.svgd_bandwidth(d, 0.0, 1.0, 1.0)
    np.testing.assert_allclose(bw["median"], g[f"{key}_median64"], rtol=1e-12)
    np.testing.assert_allclose(bw["K"].numpy(), g[f"{key}_K64"], rtol=1e-10, atol=1e-14)
    # fp32 reference is within the north-star tolerance of the fp64 oracle at these sizes
    np.testing.assert_allclose(bw["K"].numpy(), g[f"{key}_K32"], rtol=5e-5, atol=1e-6)


@pytest.mark.parametrize("n,D", RBF_CASES)
def test_fused_coefficients_reproduce_reference_phi(golden, n, D):
    """out = K G + A X must equal the reference's -(K@(-G') + s*gradK/N), G' = G + l2/2 X."""
    g = golden("rbf.npz")
    key = f"n{n}_D{D}"
    X = t(g[f"{key}_X"])
    gen = torch.Generator().manual_seed(n * 1000 + D)
    G = 1e-3 * torch.randn(X.shape, generator=gen)
    l2, s, N = 0.01, 1.7, 768.0
    out, info = O.svgd_step_fused(X, G, l2, s, N)
    K64, gK64 = t(g[f"{key}_K64"]), t(g[f"{key}_gK64"])
    phi = K64 @ (-(G.double() + l2 / 2 * X.double())) + s * gK64 / N
    np.testing.assert_allclose(out.numpy(), (-phi).numpy(), rtol=1e-9, atol=1e-13)
    # and the reference-order restatement (same ATen ops as svgd.py) agrees in fp64
    ro = O.svgd_step_reference_order(X, G, l2, s, N, dtype=torch.float64)
    np.testing.assert_allclose(ro.numpy(), (-phi).numpy(), rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("n,D", [(10, 501), (5, 37)])
def test_h_override(golden, n, D):
    g = golden("rbf.npz")
    key = f"n{n}_D{D}"
    X = t(g[f"{key}_X"])
    bw = O.svgd_bandwidth(O.svgd_pairdist(X), 0.0, 1.0, 1.0, h_override=0.7)
    assert bw["h"] == 0.7
    np.testing.assert_allclose(bw["K"].numpy(), g[f"{key}_K64_h07"], rtol=1e-10, atol=1e-300)


def test_median_rank_rule():
    """sorted position p >= n maps to pair order statistic floor((p-n)/2)."""
    for n in (2, 3, 5, 10, 20):
        gen = torch.Generator().manual_seed(n)
        X = torch.randn(n, 50, generator=gen)
        d = O.svgd_pairdist(X)
        bw = O.svgd_bandwidth(d, 0.0, 1.0, 1.0)
        q = torch.quantile(torch.sqrt(d).pow(2), 0.5).item()
        assert abs(bw["median"] - q) <= 1e-12 * max(1.0, abs(q))
        iu = torch.triu_indices(n, n, 1)
        pairs = torch.sort(d[iu[0], iu[1]]).values
        nn_ = n * n
        for which, pos in ((0, (nn_ - 1) // 2), (1, nn_ // 2)):
            i, j = divmod(bw["sel"][which], n)
            expect = 0.0 if pos < n else pairs[(pos - n) // 2].item()
            assert abs(d[i, j].item() - expect) <= 1e-12 * max(1.0, expect)


def test_svgd_reference_steps_new_grads(golden):
    """The gradients the reference hands to the base optimizer (recorded inside the real
    SVGDOptimizer.step) equal the oracle's fused step on the particles/gradients of that step."""
    g = golden("svgd_steps.npz")
    import golden_models as gm
    n, D = g["init"].shape
    parts_before = g["init"]
    for s in range(g["losses"].size):
        X = t(parts_before)
        model = gm.make_mlp()
        G = []
        for i in range(n):
            gm.load_flat(model.parameters(), parts_before[i])
            model.zero_grad()
            fwd, bwd = gm.mse_closures(model, t(g["xs"][s]), t(g["ys"][s]))
            bwd(fwd())
            G.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]))
        G = torch.stack(G)
        out, _ = O.svgd_step_fused(X, G, 0.01, 1.0, 768.0)
        np.testing.assert_allclose(out.numpy(), g["new_grads"][s], rtol=2e-5, atol=2e-6)
        parts_before = g["particles"][s]


def test_swag_matches_reference(golden):
    g = golden("swag_steps.npz")
    D = g["init"].size
    K = g["deviations"].shape[1]
    mean = t(g["init"]).clone()
    sq = mean ** 2
    ring = torch.zeros(K, D)
    dev_ref_layout = torch.zeros(D, K)
    updates = 0
    # gen_golden: start_epoch=1 after 2 steps, update_interval=2 -> steps 3,5,7,9,11 (0-based) collect
    for s in range(g["thetas"].shape[0]):
        if s >= 2 and (s - 2 + 1) % 2 == 0:
            updates += 1
            theta = t(g["thetas"][s])
            mean, sq, col = O.swag_update(theta, mean, sq, updates)
            ring[(updates - 1) % K] = col
            dev_ref_layout = O.swag_roll_deviations(dev_ref_layout, col)
    assert updates == int(g["updates"])
    np.testing.assert_array_equal(mean.numpy(), g["mean"])
    np.testing.assert_array_equal(sq.numpy(), g["sq"])
    np.testing.assert_array_equal(dev_ref_layout.numpy(), g["deviations"])
    np.testing.assert_array_equal(O.swag_ring_to_reference(ring, updates).numpy(), g["deviations"])
    sizes = g["eps_sizes"]
    assert list(sizes) == [K, D, K, D]
    off = 0
    for k in range(2):
        ek = t(g["eps"][off:off + K]); off += K
        ed = t(g["eps"][off:off + D]); off += D
        smp = O.swag_sample(mean, sq, dev_ref_layout, ek, ed)
        np.testing.assert_allclose(smp.numpy(), g["samples"][k], rtol=RTOL, atol=ATOL)


def test_ivon_matches_reference(golden):
    g = golden("ivon_steps.npz")
    import golden_models as gm
    D = g["init"].size
    S = 2
    mean = t(g["init"]).clone()
    mom = torch.zeros(D)
    prec = torch.full((D,), 10.0 / 768)
    model = gm.make_mlp()
    eps_all = g["eps"]
    sizes = [p.numel() for p in model.parameters()]
    off = 0
    for s in range(g["losses"].size):
        dsum, acc, loss_acc = None, None, 0.0
        for _ in range(S):
            eps = t(eps_all[off:off + D]); off += D
            theta, dsum = O.ivon_sample(mean, prec, dsum, eps, 768.0)
            gm.load_flat(model.parameters(), theta.numpy())
            model.zero_grad()
            fwd, bwd = gm.mse_closures(model, t(g["xs"][s]), t(g["ys"][s]))
            loss = fwd(); bwd(loss)
            loss_acc += loss.item()
            grad = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
            acc = grad if acc is None else acc + grad
        mean, mom, prec = O.ivon_update(acc, dsum, mean, mom, prec, mc_samples=S, step=s + 1, lr=1e-2, prior_prec=10.0,
                                        n_eff=768.0, damping=1e-3)
        np.testing.assert_allclose(loss_acc / S, g["losses"][s], rtol=1e-6)
        np.testing.assert_allclose(mean.numpy(), g["means"][s], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(mom.numpy(), g["momenta"][s], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(prec.numpy(), g["precisions"][s], rtol=RTOL, atol=ATOL)
    assert sum(sizes) == D


def test_gaussian_parameter_and_priors_match_reference(golden):
    g = golden("vectors.npz")
    mu, rho, eps = t(g["mu"]), t(g["rho"]), t(g["eps"])
    np.testing.assert_allclose(O.gauss_sample_fwd(mu, rho, eps).numpy(), g["w"], rtol=RTOL, atol=ATOL)
    gmu, grho = O.gauss_sample_bwd(t(g["grad_w"]), rho, eps)
    np.testing.assert_array_equal(gmu.numpy(), g["grad_mu"])
    np.testing.assert_allclose(grho.numpy(), g["grad_rho"], rtol=RTOL, atol=ATOL)
    val, kgm, kgr = O.kl_gauss(mu, rho, 0.5, 0.8)
    np.testing.assert_allclose(val.item(), g["kl_gauss"], rtol=RTOL)
    np.testing.assert_allclose(kgm.numpy(), g["kl_gauss_gmu"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(kgr.numpy(), g["kl_gauss_grho"], rtol=RTOL, atol=ATOL)
    mval, mg = O.kl_mixture(t(g["mix_mu"]), 0.3, 1.0, 0.0025)
    np.testing.assert_allclose(mval.item(), g["kl_mix"], rtol=RTOL)
    np.testing.assert_allclose(mg.numpy(), g["kl_mix_gmu"], rtol=RTOL, atol=ATOL)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from Random123 (kat_vectors): ctr/key all-zero and
    all-ones, plus the pi digits test."""
    r = O.philox4x32_10(np.array([0], dtype=np.uint64), 0, 0)[0]
    assert [hex(int(v)) for v in r] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    q = np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64)
    r = O.philox4x32_10(q, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)[0]
    assert [hex(int(v)) for v in r] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    q = np.array([(0x85a308d3 << 32) | 0x243f6a88], dtype=np.uint64)
    r = O.philox4x32_10(q, (0x299f31d0 << 32) | 0xa4093822, (0x03707344 << 32) | 0x13198a2e)[0]
    assert [hex(int(v)) for v in r] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_normal_statistics():
    z = O.philox_normal(400000, seed=1234, stream_id=7)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    # shard independence: a slice generated with elem0 equals the slice of the whole
    part = O.philox_normal(1000, seed=1234, stream_id=7, elem0=4096)
    np.testing.assert_array_equal(part, z[4096:5096])
