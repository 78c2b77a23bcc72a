"""Host-side logic of the drop-in optimizer classes, on CPU.

The classes run unmodified; the C-ABI underneath is replaced by tests/fake_abi.py (the oracle
on host pointers), and every result is compared with fixtures recorded from the UNMODIFIED
reference (oracle/gen_golden.py).  What this pins: closure protocol, particle / parameter
aliasing, shared base optimizer stepped once per particle, SWAG gating and ring buffer,
iVON MC loop, BBB loss assembly, GradScaler handling, state-dict layouts, MultiX predict.
"""
from __future__ import annotations

import copy

import numpy as np
import pytest
import torch

import fake_abi
import golden_models as gm
from conftest import ATOL, RTOL

import beyond_deep_ensembles_b200 as bde
from beyond_deep_ensembles_b200 import noise
from beyond_deep_ensembles_b200.layout import ParamLayout, shard_bounds


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture
def fake(monkeypatch):
    torch.set_num_threads(1)
    return fake_abi.install(monkeypatch)


def tape(eps_flat, sizes):
    """Replay recorded noise draws in order, whatever kind is asked for."""
    chunks, off = [], 0
    for s in sizes:
        chunks.append(eps_flat[off:off + s]); off += s
    it = iter(chunks)
    return lambda kind, numel: t(next(it))


# ---------------------------------------------------------------- layout
def test_layout_views_alias_and_roundtrip():
    params = [torch.randn(3, 5), torch.randn(7), torch.randn(2, 2, 2)]
    L = ParamLayout(params)
    assert L.logical_size == 15 + 7 + 8 and L.size % 64 == 0 and all(o % 64 == 0 for o in L.offsets)
    arena = L.new_arena(2, "cpu")
    views = L.views(arena[1])
    views[1].fill_(3.0)
    assert arena[1, L.offsets[1]:L.offsets[1] + 7].eq(3.0).all() and arena[0].eq(0).all()
    vec = torch.arange(L.logical_size, dtype=torch.float32)
    assert torch.equal(L.to_logical(L.from_logical(vec)), vec)


def test_shard_bounds_cover_and_align():
    for D in (1, 63, 64, 1000, 273610, 100_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(D, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == D
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 64 == 0 or lo == D for lo, _ in spans)  # empty tail shards start at D


# ---------------------------------------------------------------- SVGD
def build_svgd(g, base_cls=torch.optim.Adam, **base_kw):
    n, D = g["init"].shape
    model = gm.make_mlp()
    gm.load_flat(model.parameters(), g["init"][0])
    k = {"k": 0}

    def reset():
        k["k"] += 1
        gm.load_flat(model.parameters(), g["init"][k["k"]])

    base = base_cls(model.parameters(), **(base_kw or {"lr": 1e-2}))
    opt = bde.SVGDOptimizer(model.parameters(), reset, base, particle_count=n, dataset_size=768, l2_reg=0.01,
                            kernel_grad_scale=1.0)
    assert k["k"] == n - 1
    return model, opt


def test_svgd_steps_match_reference(fake, golden):
    g = golden("svgd_steps.npz")
    n, D = g["init"].shape
    model, opt = build_svgd(g)
    assert len(opt.param_groups) == 4  # one group per tensor (svgd.py:50)
    for s in range(g["losses"].size):
        fwd, bwd = gm.mse_closures(model, t(g["xs"][s]), t(g["ys"][s]))
        loss = opt.step(fwd, bwd)
        np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
        parts = np.stack([gm.flat_params(opt._params_for_particle(i)) for i in range(n)])
        np.testing.assert_allclose(parts, g["particles"][s], rtol=2e-5, atol=2e-6)
        # the model's parameters alias the LAST particle after a step (svgd.py:96)
        assert all(p.data_ptr() == v.data_ptr() for p, v in zip(model.parameters(), opt._params_for_particle(n - 1)))
    # cursor semantics of sample_parameters (svgd.py:107-112)
    for k in range(n + 2):
        opt.sample_parameters()
        np.testing.assert_allclose(gm.flat_params(model.parameters()), g["sampled"][k], rtol=2e-5, atol=2e-6)
    assert fake.calls.count("pairdist") == g["losses"].size and fake.calls.count("apply") == g["losses"].size


def test_svgd_state_dict_roundtrip_and_keys(fake, golden):
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(g)
    fwd, bwd = gm.mse_closures(model, t(g["xs"][0]), t(g["ys"][0]))
    opt.step(fwd, bwd)
    sd = copy.deepcopy(opt.state_dict())
    assert {"__base_optimizer", "__l2_reg", "__dataset_size", "__current_particle", "__particle_count",
            "__kernel_grad_scale", 0, 1, 2, 3} <= set(sd["state"].keys())
    assert set(sd["state"][0].keys()) == {f"particle_{i}" for i in range(10)}
    assert sd["state"][0]["particle_3"].shape == (50, 8)
    model2, opt2 = build_svgd(g)
    opt2.load_state_dict(sd)
    for i in range(10):
        a = gm.flat_params(opt._params_for_particle(i)); b = gm.flat_params(opt2._params_for_particle(i))
        np.testing.assert_array_equal(a, b)
    # loaded particles live in the arena again
    assert opt2.state[next(iter(opt2._params()))]["particle_0"].data_ptr() == opt2._xviews[0][0].data_ptr()


def test_svgd_grad_scaler_protocol(fake, golden):
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(g)
    scaler = torch.amp.GradScaler("cpu", init_scale=1024.0)
    opt.init_grad_scaler(scaler)
    x, y = t(g["xs"][0]), t(g["ys"][0])

    def fwd():
        return ((model(x).squeeze(-1) - y) ** 2).mean()

    def bwd(loss):
        scaler.scale(loss).backward()

    loss = opt.step(fwd, bwd, grad_scaler=scaler)
    scaler.update()
    np.testing.assert_allclose(loss.item(), g["losses"][0], rtol=1e-5)
    parts = np.stack([gm.flat_params(opt._params_for_particle(i)) for i in range(10)])
    np.testing.assert_allclose(parts, g["particles"][0], rtol=5e-5, atol=5e-6)
    assert "found_inf_per_device" in opt.state  # the reference leaves this key behind (algo.py:73)


def test_rbf_function(fake, golden):
    g = golden("rbf.npz")
    X = t(g["n10_D501_X"])
    K, gK = bde.rbf(X)
    np.testing.assert_allclose(K.numpy(), g["n10_D501_K64"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(gK.numpy(), g["n10_D501_gK64"], rtol=RTOL, atol=ATOL)


# ---------------------------------------------------------------- SWAG
def build_swag(g, K=4):
    model = gm.make_mlp()
    gm.load_flat(model.parameters(), g["init"])
    base = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    opt = bde.SwagOptimizer(model.parameters(), base, update_interval=2, start_epoch=1, deviation_samples=K)
    return model, opt


def run_swag(model, opt, g):
    for s in range(g["thetas"].shape[0]):
        if s == 2:
            opt.complete_epoch()
        fwd, bwd = gm.mse_closures(model, t(g["xs"][s]), t(g["ys"][s]))
        loss = opt.step(fwd, bwd)
        np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
        np.testing.assert_allclose(gm.flat_params(model.parameters()), g["thetas"][s], rtol=RTOL, atol=ATOL)


def test_swag_matches_reference(fake, golden):
    g = golden("swag_steps.npz")
    model, opt = build_swag(g)
    run_swag(model, opt, g)
    assert opt.state["__updates"] == int(g["updates"]) == fake.calls.count("swag_update")
    sd = opt.state_dict()
    st = sd["state"]
    np.testing.assert_allclose(st["__mean"].numpy(), g["mean"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st["__sq_weights"].numpy(), g["sq"], rtol=RTOL, atol=ATOL)
    assert st["__deviations"].shape == g["deviations"].shape  # [D, K], roll order
    np.testing.assert_allclose(st["__deviations"].numpy(), g["deviations"], rtol=RTOL, atol=ATOL)
    assert "__mean" not in opt.state  # export only
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        for k in range(2):
            opt.sample_parameters()
            np.testing.assert_allclose(gm.flat_params(model.parameters()), g["samples"][k], rtol=RTOL, atol=ATOL)
    assert opt.state["__params_dirty"]
    fwd, bwd = gm.mse_closures(model, t(g["xs"][0]), t(g["ys"][0]))
    loss = opt.step(fwd, bwd)  # restores the training weights first (swag.py:38)
    np.testing.assert_allclose(loss.item(), g["loss_after"], rtol=1e-5)
    np.testing.assert_allclose(gm.flat_params(model.parameters()), g["theta_after"], rtol=RTOL, atol=ATOL)


def test_swag_loads_reference_layout_checkpoint(fake, golden):
    g = golden("swag_steps.npz")
    model, opt = build_swag(g)
    run_swag(model, opt, g)
    sd = copy.deepcopy(opt.state_dict())
    model2, opt2 = build_swag(g)
    model2.load_state_dict(model.state_dict())
    opt2.load_state_dict(sd)
    assert opt2.state["__updates"] == int(g["updates"])
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        opt2.sample_parameters()
    np.testing.assert_allclose(gm.flat_params(model2.parameters()), g["samples"][0], rtol=RTOL, atol=ATOL)


# ---------------------------------------------------------------- iVON
def build_ivon(g):
    model = gm.make_mlp()
    gm.load_flat(model.parameters(), g["init"])
    opt = bde.iVONOptimizer(model.parameters(), lr=1e-2, prior_prec=10.0, dataset_size=768, damping=1e-3,
                            mc_samples=2, augmentation=1.0, tempering=1.0)
    return model, opt


def ivon_tape(g, model):
    """The reference draws per tensor; this repo draws once per arena: regroup the tape."""
    sizes = [p.numel() for p in model.parameters()]
    per_call = sum(sizes)
    n_calls = g["eps"].size // per_call
    return tape(g["eps"], [per_call] * n_calls)


def test_ivon_matches_reference(fake, golden):
    g = golden("ivon_steps.npz")
    model, opt = build_ivon(g)
    assert opt.get_base_optimizer() is opt
    with noise.inject(ivon_tape(g, model)):
        for s in range(g["losses"].size):
            fwd, bwd = gm.mse_closures(model, t(g["xs"][s]), t(g["ys"][s]))
            loss = opt.step(fwd, bwd)
            np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
            st = [opt.state[p] for p in model.parameters()]
            for name, key in (("mean", "means"), ("momentum", "momenta"), ("precision", "precisions")):
                got = torch.cat([x[name].reshape(-1) for x in st]).numpy()
                np.testing.assert_allclose(got, g[key][s], rtol=RTOL, atol=ATOL)
            assert st[0]["delta"] is not None and st[0]["acc_grad"] is not None
        opt.sample_parameters()
        np.testing.assert_allclose(gm.flat_params(model.parameters()), g["sampled"], rtol=RTOL, atol=ATOL)
    assert opt.param_groups[0]["step"] == g["losses"].size


def test_ivon_state_dict_roundtrip(fake, golden):
    g = golden("ivon_steps.npz")
    model, opt = build_ivon(g)
    with noise.inject(ivon_tape(g, model)):
        fwd, bwd = gm.mse_closures(model, t(g["xs"][0]), t(g["ys"][0]))
        opt.step(fwd, bwd)
    sd = copy.deepcopy(opt.state_dict())
    assert set(sd["state"][0].keys()) == {"mean", "momentum", "precision", "delta", "acc_grad"}
    assert sd["param_groups"][0]["step"] == 1 and sd["param_groups"][0]["lr"] == 1e-2
    model2, opt2 = build_ivon(g)
    opt2.load_state_dict(sd)
    for p, q in zip(model.parameters(), model2.parameters()):
        for name in ("mean", "momentum", "precision"):
            assert torch.equal(opt.state[p][name], opt2.state[q][name])
    assert opt2.state[next(iter(model2.parameters()))]["mean"].data_ptr() == opt2._arenas[0]["views"]["mean"][0].data_ptr()
    # lr schedulers act on the optimizer itself (ivorn.py:117-118)
    sched = torch.optim.lr_scheduler.LambdaLR(opt2.get_base_optimizer(), lambda e: 0.5)
    assert opt2.param_groups[0]["lr"] == 0.5e-2


# ---------------------------------------------------------------- BBB / Rank-1
def test_bbb_rank1_matches_reference(fake, golden):
    g = golden("bbb_steps.npz")
    model = gm.Rank1MLP(bde.GaussianParameter)
    init = {k[len("init/"):]: g[k] for k in g.files if k.startswith("init/")}
    gm.init_rank1(model, init)
    prior = bde.GaussianPrior(0.5, 0.8)
    base = torch.optim.Adam(model.parameters(), lr=1e-2)
    opt = bde.BBBOptimizer(model.parameters(), base, prior, dataset_size=100, mc_samples=2, kl_rescaling=0.5,
                           components=1, l2_scale=0.01)
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        for s in range(g["losses"].size):
            fwd, bwd = gm.mse_closures(model, t(g["xs"][s]), t(g["ys"][s]))
            loss = opt.step(fwd, bwd)
            np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
            for name, p in model.named_parameters():
                np.testing.assert_allclose(p.detach().numpy(), g[f"step{s}/{name}"], rtol=5e-5, atol=5e-6,
                                           err_msg=f"{name} after step {s}")
    assert fake.calls.count("kl_gauss") == 2 * 4 * g["losses"].size  # value + grad, 4 Gaussian tensors
    assert opt.sample_parameters() is None


def test_bbb_skips_step_on_nan_loss(fake, golden):
    model = gm.Rank1MLP(bde.GaussianParameter)
    for p in model.parameters():
        torch.nn.init.constant_(p, 0.1)
    base = torch.optim.SGD(model.parameters(), lr=0.1)
    opt = bde.BBBOptimizer(model.parameters(), base, bde.GaussianPrior(0.0, 1.0), dataset_size=10)
    before = gm.flat_params(model.parameters())
    called = {"bwd": 0}
    loss = opt.step(lambda: torch.tensor(float("nan")), lambda l: called.__setitem__("bwd", called["bwd"] + 1))
    assert loss.isnan() and called["bwd"] == 0
    np.testing.assert_array_equal(before, gm.flat_params(model.parameters()))


def test_mixture_prior_kl_through_gaussian_parameter(fake, golden):
    g = golden("vectors.npz")
    gp = bde.GaussianParameter(g["mix_mu"].size)
    with torch.no_grad():
        gp.mean.copy_(t(g["mix_mu"])); gp.rho.copy_(t(g["rho"]))
    kl = gp.mean.get_parameter_kl(bde.MixturePrior(0.3, 1.0, 0.0025))
    kl.backward()
    np.testing.assert_allclose(kl.item(), g["kl_mix"], rtol=RTOL)
    np.testing.assert_allclose(gp.mean.grad.numpy(), g["kl_mix_gmu"], rtol=RTOL, atol=ATOL)


# ---------------------------------------------------------------- wrappers
def test_last_layer_wrapper_accumulates_deterministic_grads(fake, golden):
    g = golden("ivon_steps.npz")
    body = torch.nn.Linear(8, 8)
    head = gm.make_mlp()
    ll = bde.iVONOptimizer(head.parameters(), lr=1e-2, prior_prec=10.0, dataset_size=768, mc_samples=3)
    det = torch.optim.SGD(body.parameters(), lr=0.1)
    opt = bde.LastLayerBayesianOptimizer(ll, det)
    x, y = t(g["xs"][0]), t(g["ys"][0])
    seen = []

    def fwd():
        return ((head(body(x)).squeeze(-1) - y) ** 2).mean()

    def bwd(loss):
        loss.backward()
        seen.append(body.weight.grad.clone())

    w0 = body.weight.detach().clone()
    opt.step(fwd, bwd)
    # gradients of the deterministic body accumulate over the 3 MC passes (algo.py:100-103)
    assert len(seen) == 3 and not torch.allclose(seen[0], seen[2])
    torch.testing.assert_close(body.weight.detach(), w0 - 0.1 * seen[2])
    with pytest.raises(ValueError):
        opt.step(fwd, bwd, grad_scaler=torch.amp.GradScaler("cpu"))
    with pytest.raises(RuntimeError):
        opt.get_base_optimizer()
    assert set(opt.state_dict().keys()) == {"ll_bayesian_optimizer", "deterministic_optimizer"}


def test_deep_ensemble_predict_matches_reference(fake, golden):
    g = golden("ensemble_predict.npz")
    pairs = []
    for m in range(g["inits"].shape[0]):
        init = g["inits"][m]
        model = gm.make_mlp()
        gm.load_flat(model.parameters(), init[0])
        k = {"k": 0}

        def reset(model=model, init=init, k=k):
            k["k"] += 1
            gm.load_flat(model.parameters(), init[k["k"]])

        opt = bde.SVGDOptimizer(model.parameters(), reset, torch.optim.SGD(model.parameters(), lr=0.1),
                                particle_count=init.shape[0], dataset_size=100)
        pairs.append((model, opt))
    ens = bde.DeepEnsemble(pairs)
    with torch.no_grad():
        preds = ens.predict(lambda mdl: mdl(t(g["x"])).squeeze(-1), samples=7)
    np.testing.assert_allclose(preds.numpy(), g["preds"], rtol=1e-6, atol=1e-7)
    sd = ens.state_dict()
    assert set(sd.keys()) == {"models", "optimizers"} and len(sd["optimizers"]) == 2
    ens.load_state_dict(copy.deepcopy(sd))
