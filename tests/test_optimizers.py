"""The drop-in optimizer classes against fixtures recorded from the UNMODIFIED reference
(oracle/gen_golden.py), in two environments:

  oracle-abi (CPU, `-m "not gpu"`): the classes run unmodified, the C-ABI underneath is
      replaced by tests/fake_abi.py (the oracle on host pointers) — pins the host logic:
      closure protocol, particle / parameter aliasing, shared base optimizer stepped once per
      particle, SWAG gating and ring buffer, iVON MC loop, BBB loss assembly, GradScaler
      handling, state-dict layouts, MultiX predict;
  cuda (`-m gpu`): the same bodies on the real library — parity of the CUDA path.

Tolerance: the north-star fp32 rtol 1e-5 / atol 1e-6 per update; a few multiples of it where
several optimizer steps compound.
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


class Env:
    """Where a test body runs: "cpu" = oracle-backed ABI double, "cuda" = the real library."""

    def __init__(self, dev, fake):
        self.dev, self.fake = dev, fake

    def t(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def calls(self, name):
        return None if self.fake is None else self.fake.calls.count(name)


@pytest.fixture(params=[pytest.param("cpu", id="oracle-abi"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def env(request, monkeypatch):
    if request.param == "cuda":
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        from beyond_deep_ensembles_b200 import _lib
        _lib.get()  # fail loudly if the library is missing
        return Env("cuda", None)
    torch.set_num_threads(1)
    return Env("cpu", fake_abi.install(monkeypatch))


def tape(eps_flat, sizes):
    """Replay recorded noise draws in order (noise.draw moves them to the right device)."""
    chunks, off = [], 0
    for s in sizes:
        chunks.append(torch.from_numpy(np.ascontiguousarray(eps_flat[off:off + s])))
        off += s
    it = iter(chunks)
    return lambda kind, numel: next(it)


def flat(params):
    return gm.flat_params(params)


# ---------------------------------------------------------------- layout (host only)
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
def build_svgd(env, g, base_cls=torch.optim.Adam, **base_kw):
    n, D = g["init"].shape
    model = gm.make_mlp().to(env.dev)
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


def test_svgd_zero_grad_shortcut_only_when_base_owns_exactly_these_params(env, golden):
    """svgd.py:74 calls base_optimizer.zero_grad() per particle.  The gather path folds it into the rebinding loop
    (grad = None) only when the base optimizer holds exactly the SVGD parameters and its zero_grad is torch's own;
    a base optimizer with an extra parameter group, or an overridden zero_grad, is called as the reference calls it."""
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(env, g)
    plist = list(model.parameters())
    base = opt.get_base_optimizer()
    assert opt._base_owns_exactly(base, plist)
    extra = torch.nn.Parameter(torch.zeros(3, device=env.dev))
    base.add_param_group({"params": [extra]})
    opt._base_zero_fast = None
    assert not opt._base_owns_exactly(base, plist)
    opt.prebind_grads = False
    extra.grad = torch.ones_like(extra)
    fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
    opt.step(fwd, bwd)
    assert extra.grad is None            # base.zero_grad() ran (the shortcut would have left it alone)

    class Hooked(torch.optim.SGD):
        def zero_grad(self, set_to_none=True):
            self.calls = getattr(self, "calls", 0) + 1
            super().zero_grad(set_to_none)
    model2, opt2 = build_svgd(env, g, base_cls=Hooked)
    opt2.prebind_grads = False
    fwd, bwd = gm.mse_closures(model2, env.t(g["xs"][0]), env.t(g["ys"][0]))
    opt2.step(fwd, bwd)
    assert opt2.get_base_optimizer().calls == g["init"].shape[0]


@pytest.mark.parametrize("capture", ["prebound", "gather", "closure-zeroes-grads"])
def test_svgd_gradient_capture_forms_match_reference(env, golden, capture):
    """the particles' gradients reach the G arena through pre-bound `.grad` views (autograd accumulates
    into the zeroed arena row: no gather launch), through the gather launch, or — when the closure clears the
    gradients itself — through the gather as a fallback; all three reproduce the reference's steps."""
    g = golden("svgd_steps.npz")
    n, D = g["init"].shape
    model, opt = build_svgd(env, g)
    opt.prebind_grads = capture != "gather"
    for s in range(g["losses"].size):
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
        if capture == "closure-zeroes-grads":
            inner = fwd

            def fwd(inner=inner):
                model.zero_grad(set_to_none=True)
                return inner()
        n_gathers = env.calls("mtc")
        loss = opt.step(fwd, bwd)
        if env.fake is not None:
            assert env.calls("mtc") - n_gathers == (0 if capture == "prebound" else n)
        np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
        parts = np.stack([flat(opt._params_for_particle(i)) for i in range(n)])
        np.testing.assert_allclose(parts, g["particles"][s], rtol=3e-5, atol=3e-6)


@pytest.mark.parametrize("mode", ["train-step", "fused-base", "base-step"])
def test_svgd_steps_match_reference(env, golden, mode):
    """Three reference steps with a shared Adam base optimizer (10 Adam steps per SVGD step); fused-base = the
    base-optimizer steps run inside the apply kernel (f1), train-step = that launch also produces the next
    step's pair distances (K1 runs once, on the first step only), base-step = base.step() per particle."""
    g = golden("svgd_steps.npz")
    n, D = g["init"].shape
    model, opt = build_svgd(env, g)
    fused = mode != "base-step"
    opt.fuse_base_optimizer = fused
    opt.reuse_pair_distances = mode == "train-step"
    base = opt.get_base_optimizer()
    assert len(opt.param_groups) == 4  # one group per tensor (svgd.py:50)
    for s in range(g["losses"].size):
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
        loss = opt.step(fwd, bwd)
        np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
        parts = np.stack([flat(opt._params_for_particle(i)) for i in range(n)])
        np.testing.assert_allclose(parts, g["particles"][s], rtol=3e-5, atol=3e-6)
        # the model's parameters alias the LAST particle after a step (svgd.py:96) and carry its new gradient (:94)
        assert all(p.data_ptr() == v.data_ptr() for p, v in zip(model.parameters(), opt._params_for_particle(n - 1)))
        np.testing.assert_allclose(flat(p.grad for p in model.parameters()), g["new_grads"][s][n - 1], rtol=1e-4, atol=ATOL)
        # the shared optimizer has stepped once per particle
        assert all(float(base.state[p]["step"]) == n * (s + 1) for p in model.parameters())
    # cursor semantics of sample_parameters (svgd.py:107-112)
    for k in range(n + 2):
        opt.sample_parameters()
        np.testing.assert_allclose(flat(model.parameters()), g["sampled"][k], rtol=3e-5, atol=3e-6)
    if env.fake:
        steps = g["losses"].size
        assert env.calls("pairdist") == (1 if mode == "train-step" else steps)
        assert env.calls("train_step_adam") == (steps if mode == "train-step" else 0)
        assert env.calls("apply_adam") == (steps if mode == "fused-base" else 0)
        assert env.calls("apply") == (0 if fused else steps)


@pytest.mark.parametrize("fused", [True, False], ids=["fused-base", "base-step"])
def test_svgd_sgd_nesterov_schedule_and_checkpoint_match_reference(env, golden, fused):
    """CIFAR base optimizer (SGD momentum 0.9, Nesterov, weight decay) with a StepLR schedule; after step 2
    the base optimizer is checkpointed and restored into a FRESH optimizer object (state tensors no longer
    alias the arena) — the run must still follow the reference's trajectory (svgd.py:92-103)."""
    g = golden("svgd_sgd_steps.npz")
    n, D = g["init"].shape
    kw = dict(lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-4)
    model = gm.make_mlp().to(env.dev)
    gm.load_flat(model.parameters(), g["init"][0])
    k = {"k": 0}

    def reset():
        k["k"] += 1
        gm.load_flat(model.parameters(), g["init"][k["k"]])

    base = torch.optim.SGD(model.parameters(), **kw)
    opt = bde.SVGDOptimizer(model.parameters(), reset, base, particle_count=n, dataset_size=768, l2_reg=3e-4,
                            kernel_grad_scale=1.0)
    opt.fuse_base_optimizer = fused
    sched = torch.optim.lr_scheduler.StepLR(base, step_size=2, gamma=0.5)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # e.g. "lr_scheduler.step() before optimizer.step()"
        for s in range(g["losses"].size):
            if s == 2:  # checkpoint round trip of the caller's base optimizer
                sd = copy.deepcopy(base.state_dict())
                assert set(sd["state"][0].keys()) == {"momentum_buffer"}
                base.load_state_dict(sd)
            assert base.param_groups[0]["lr"] == pytest.approx(g["lrs"][s])
            fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
            loss = opt.step(fwd, bwd)
            sched.step()
            np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
            parts = np.stack([flat(opt._params_for_particle(i)) for i in range(n)])
            np.testing.assert_allclose(parts, g["particles"][s], rtol=3e-5, atol=3e-6)
            bufs = flat(base.state[p]["momentum_buffer"] for p in model.parameters())
            np.testing.assert_allclose(bufs, g["momentum_buffers"][s], rtol=1e-4, atol=1e-6)
    if env.fake:  # default reuse_pair_distances: the training-step launch; K1 only on the first step
        assert env.calls("train_step_sgd") == (g["losses"].size if fused else 0)
        assert env.calls("pairdist") == (1 if fused else g["losses"].size)


def test_svgd_kernel_cache_invalidation(env, golden):
    """The pair distances cached by the training-step launch are dropped by load_state_dict(), by
    invalidate_kernel_cache(), by a change of l2_reg / dataset_size, and never used by an unfused step."""
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(env, g)
    fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
    opt.step(fwd, bwd)
    assert opt._cached == "kernel"
    opt.load_state_dict(opt.state_dict())
    assert opt._cached is None
    opt.step(fwd, bwd)
    assert opt._cached == "kernel"
    opt.invalidate_kernel_cache()
    assert opt._cached is None
    opt.step(fwd, bwd)
    opt.state["__l2_reg"] = 0.5       # hyper-parameters enter K1b: the cached K / A no longer apply
    before = env.calls("pairdist")
    opt.step(fwd, bwd)
    if env.fake:
        assert env.calls("pairdist") == before + 1
    opt.fuse_base_optimizer = False   # base.step() moves the particles outside the kernels
    opt.step(fwd, bwd)
    assert opt._cached is None
    # two column segments (per-group hyper-parameters): no single-pass form, K1 runs every step
    plist = list(model.parameters())
    opt.fuse_base_optimizer = True
    opt.state["__base_optimizer"] = torch.optim.SGD([{"params": plist[:2], "lr": 0.1}, {"params": plist[2:], "lr": 0.01}], lr=1.0)
    opt.step(fwd, bwd)
    assert opt._cached is None


def test_svgd_fused_plan_recognition(env, golden):
    """What is fused and what goes through base.step(): stock SGD/Adam/AdamW only, no scaler, no hooks."""
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(env, g)
    plist = list(model.parameters())

    def plan(base, scaler=None):
        opt._fused_plan = None
        return opt._plan_for(base, scaler, plist)

    assert plan(torch.optim.SGD(plist, lr=0.1, momentum=0.9)).kind == "sgd"
    assert plan(torch.optim.Adam(plist, lr=0.1)).kind == "adam"
    assert plan(torch.optim.AdamW(plist, lr=0.1)).kind == "adamw"
    assert plan(torch.optim.Adam(plist, lr=0.1, amsgrad=True)) is None
    assert plan(torch.optim.SGD(plist, lr=0.1, maximize=True)) is None
    assert plan(torch.optim.RMSprop(plist, lr=0.1)) is None
    assert plan(torch.optim.SGD(plist[:2], lr=0.1)) is None  # does not cover every particle tensor

    class MySGD(torch.optim.SGD):
        pass

    assert plan(MySGD(plist, lr=0.1)) is None  # subclasses may override step()
    hooked = torch.optim.SGD(plist, lr=0.1)
    hooked.register_step_post_hook(lambda *a: None)
    assert plan(hooked) is None
    # two groups with different hyper-parameters -> two column segments
    two = torch.optim.SGD([{"params": plist[:2], "lr": 0.1}, {"params": plist[2:], "lr": 0.01, "weight_decay": 0.1}], lr=1.0)
    p2 = plan(two)
    assert [(c0, c1) for c0, c1, _ in p2.segments] == [(0, opt._layout.offsets[2]), (opt._layout.offsets[2], opt._layout.size)]
    opt.fuse_base_optimizer = False
    assert plan(torch.optim.SGD(plist, lr=0.1)) is None


def test_svgd_fused_two_param_groups_equal_base_step(env, golden):
    """Per-group hyper-parameters: the fused path (one launch per column segment) against base.step()."""
    g = golden("svgd_steps.npz")
    results = []
    for fused in (True, False):
        model, opt = build_svgd(env, g)
        plist = list(model.parameters())
        base = torch.optim.AdamW([{"params": plist[:2], "lr": 1e-2, "weight_decay": 0.05},
                                  {"params": plist[2:], "lr": 3e-3, "betas": (0.8, 0.99)}], lr=1.0)
        opt.state["__base_optimizer"] = base
        opt.fuse_base_optimizer = fused
        for s in range(2):
            fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
            opt.step(fwd, bwd)
        results.append(np.stack([flat(opt._params_for_particle(i)) for i in range(10)]))
    np.testing.assert_allclose(results[0], results[1], rtol=3e-5, atol=3e-6)


@pytest.mark.parametrize("fused", [True, False], ids=["fused-base", "base-step"])
def test_svgd_new_gradients_match_reference_first_step(env, golden, fused):
    """Tight check of one posterior update: the gradients handed to the base optimizer."""
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(env, g, base_cls=torch.optim.SGD, lr=0.0)
    opt.fuse_base_optimizer = fused
    fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
    opt.step(fwd, bwd)
    if fused:  # only the last particle's new gradient leaves the kernel (what svgd.py:94 leaves in param.grad)
        out = opt._layout.to_logical(opt._out_last).cpu().numpy()
        np.testing.assert_allclose(out[0], g["new_grads"][0][-1], rtol=RTOL, atol=ATOL)
        assert opt._out_full is None
    else:
        out = opt._layout.to_logical(opt._out).cpu().numpy()
        np.testing.assert_allclose(out, g["new_grads"][0], rtol=RTOL, atol=ATOL)


def test_svgd_state_dict_roundtrip_and_keys(env, golden):
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(env, g)
    fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
    opt.step(fwd, bwd)
    sd = copy.deepcopy(opt.state_dict())
    assert {"__base_optimizer", "__l2_reg", "__dataset_size", "__current_particle", "__particle_count",
            "__kernel_grad_scale", 0, 1, 2, 3} <= set(sd["state"].keys())
    assert set(sd["state"][0].keys()) == {f"particle_{i}" for i in range(10)}
    assert sd["state"][0]["particle_3"].shape == (50, 8)
    model2, opt2 = build_svgd(env, g)
    opt2.load_state_dict(sd)
    for i in range(10):
        np.testing.assert_array_equal(flat(opt._params_for_particle(i)), flat(opt2._params_for_particle(i)))
    # loaded particles live in the arena again
    assert opt2.state[next(iter(opt2._params()))]["particle_0"].data_ptr() == opt2._xviews[0][0].data_ptr()


def test_svgd_grad_scaler_protocol(env, golden):
    g = golden("svgd_steps.npz")
    model, opt = build_svgd(env, g)
    scaler = torch.amp.GradScaler(env.dev, init_scale=1024.0)
    opt.init_grad_scaler(scaler)
    x, y = env.t(g["xs"][0]), env.t(g["ys"][0])

    def fwd():
        return ((model(x).squeeze(-1) - y) ** 2).mean()

    def bwd(loss):
        scaler.scale(loss).backward()

    loss = opt.step(fwd, bwd, grad_scaler=scaler)
    scaler.update()
    np.testing.assert_allclose(loss.item(), g["losses"][0], rtol=1e-5)
    parts = np.stack([flat(opt._params_for_particle(i)) for i in range(10)])
    np.testing.assert_allclose(parts, g["particles"][0], rtol=5e-5, atol=5e-6)
    assert "found_inf_per_device" in opt.state  # the reference leaves this key behind (algo.py:73)


@pytest.mark.parametrize("poison", [False, True], ids=["finite", "inf-grad"])
@pytest.mark.parametrize("algo", ["svgd", "ivon"])
def test_fused_unscale_gather_equals_unscale_then_gather(env, golden, algo, poison):
    """with an active GradScaler the gradient gather also does what GradScaler.unscale_ does.  Two
    steps + scaler.update() with the fused gather and with the literal unscale_()-then-gather order give identical
    states, scaler scales and skip decisions — also when a gradient is inf (the step of that pass is skipped by the
    scaler exactly as in the reference)."""
    g = golden("svgd_steps.npz" if algo == "svgd" else "ivon_steps.npz")

    def run(fused):
        if algo == "svgd":
            model, opt = build_svgd(env, g)
        else:
            model, opt = build_ivon(env, g)
        opt.fuse_unscale_into_gather = fused
        scaler = torch.amp.GradScaler(env.dev, init_scale=1024.0, growth_interval=1)
        opt.init_grad_scaler(scaler)
        noise.set_seed(21)
        calls = {"n": 0}
        for s in range(2):
            x, y = env.t(g["xs"][s]), env.t(g["ys"][s])

            def fwd():
                return ((model(x).squeeze(-1) - y) ** 2).mean()

            def bwd(loss):
                scaler.scale(loss).backward()
                calls["n"] += 1
                if poison and calls["n"] == 2:
                    next(iter(model.parameters())).grad.view(-1)[0] = float("inf")

            opt.step(fwd, bwd, grad_scaler=scaler)
            scaler.update()
        noise.set_seed(None)
        if algo == "svgd":
            state = np.stack([flat(opt._params_for_particle(i)) for i in range(10)])
        else:
            st = [opt.state[p] for p in model.parameters()]
            state = np.concatenate([torch.cat([x_[k].reshape(-1) for x_ in st]).cpu().numpy()
                                    for k in ("mean", "momentum", "precision")])
        return state, float(scaler.get_scale())

    a, scale_a = run(True)
    if env.fake:
        assert env.calls("mtc_unscale") > 0 and env.calls("mtc") == 0
    b, scale_b = run(False)
    if env.fake:
        assert env.calls("mtc") > 0
    np.testing.assert_array_equal(a, b)
    assert scale_a == scale_b


def test_rbf_function(env, golden):
    g = golden("rbf.npz")
    X = env.t(g["n10_D501_X"])
    K, gK = bde.rbf(X)
    np.testing.assert_allclose(K.cpu().numpy(), g["n10_D501_K64"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(gK.cpu().numpy(), g["n10_D501_gK64"], rtol=RTOL, atol=ATOL)
    K2, gK2 = bde.rbf(X, h_override=0.7)
    np.testing.assert_allclose(K2.cpu().numpy(), g["n10_D501_K64_h07"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(gK2.cpu().numpy(), g["n10_D501_gK64_h07"], rtol=RTOL, atol=ATOL)


# ---------------------------------------------------------------- SWAG
def build_swag(env, g, K=4):
    model = gm.make_mlp().to(env.dev)
    gm.load_flat(model.parameters(), g["init"])
    base = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    opt = bde.SwagOptimizer(model.parameters(), base, update_interval=2, start_epoch=1, deviation_samples=K)
    return model, opt


def run_swag(env, model, opt, g):
    for s in range(g["thetas"].shape[0]):
        if s == 2:
            opt.complete_epoch()
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
        loss = opt.step(fwd, bwd)
        np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
        np.testing.assert_allclose(flat(model.parameters()), g["thetas"][s], rtol=3e-5, atol=3e-6)


def test_swag_matches_reference(env, golden):
    g = golden("swag_steps.npz")
    model, opt = build_swag(env, g)
    run_swag(env, model, opt, g)
    assert opt.state["__updates"] == int(g["updates"])
    if env.fake:
        assert env.calls("swag_update") == int(g["updates"])
    sd = opt.state_dict()
    st = sd["state"]
    np.testing.assert_allclose(st["__mean"].numpy(), g["mean"], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(st["__sq_weights"].numpy(), g["sq"], rtol=3e-5, atol=3e-6)
    assert st["__deviations"].shape == g["deviations"].shape  # [D, K] on the host, roll order
    assert st["__deviations"].device.type == "cpu"
    np.testing.assert_allclose(st["__deviations"].numpy(), g["deviations"], rtol=3e-5, atol=3e-6)
    assert "__mean" not in opt.state  # export only
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        for k in range(2):
            opt.sample_parameters()
            np.testing.assert_allclose(flat(model.parameters()), g["samples"][k], rtol=3e-5, atol=3e-6)
    assert opt.state["__params_dirty"]
    fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
    loss = opt.step(fwd, bwd)  # restores the training weights first (swag.py:38)
    np.testing.assert_allclose(loss.item(), g["loss_after"], rtol=1e-5)
    np.testing.assert_allclose(flat(model.parameters()), g["theta_after"], rtol=3e-5, atol=3e-6)


def test_swag_sample_from_reference_moments(env, golden):
    """Tight check of K4 alone: load the reference's own moments, draw with its noise."""
    g = golden("swag_steps.npz")
    model, opt = build_swag(env, g)
    sd = opt.state_dict()
    sd["state"]["__mean"] = torch.from_numpy(g["mean"])
    sd["state"]["__sq_weights"] = torch.from_numpy(g["sq"])
    sd["state"]["__deviations"] = torch.from_numpy(g["deviations"])
    sd["state"]["__updates"] = int(g["updates"])
    opt.load_state_dict(sd)
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        for k in range(2):
            opt.sample_parameters()
            np.testing.assert_allclose(flat(model.parameters()), g["samples"][k], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("count,max_rows", [(2, 99), (5, 99), (5, 2)])
def test_swag_presample_equals_sequential_draws(env, golden, count, max_rows):
    """presample(count) hands out exactly the draws that `count` single sample_parameters()
    calls produce — with injected noise (the reference's draw order) and with Philox streams — also when the
    buffer cap splits the request into several batches, and the batch is dropped by step()."""
    g = golden("swag_steps.npz")

    def loaded():
        model, opt = build_swag(env, g)
        sd = opt.state_dict()
        sd["state"]["__mean"] = torch.from_numpy(g["mean"])
        sd["state"]["__sq_weights"] = torch.from_numpy(g["sq"])
        sd["state"]["__deviations"] = torch.from_numpy(g["deviations"])
        sd["state"]["__updates"] = int(g["updates"])
        opt.load_state_dict(sd)
        return model, opt

    K, D = g["deviations"].shape[1], g["mean"].shape[0]
    gen = torch.Generator().manual_seed(3)
    zs = [torch.randn(K if i % 2 == 0 else D, generator=gen) for i in range(2 * count)]

    def draws(batched, injected):
        model, opt = loaded()
        if batched:
            opt.presample_max_bytes = 4 * opt._theta.numel() * max_rows
            opt.presample(count)
        noise.set_seed(1234)
        it = iter(zs)
        ctx = noise.inject(lambda kind, numel: next(it)) if injected else noise.inject(lambda kind, numel: None)
        out = []
        with ctx:
            for _ in range(count):
                opt.sample_parameters()
                out.append(flat(model.parameters()).copy())
        noise.set_seed(None)
        return out, model, opt

    for injected in (True, False):
        single, _, _ = draws(False, injected)
        batch, model, opt = draws(True, injected)
        for a, b in zip(single, batch):
            np.testing.assert_array_equal(a, b)
        assert any(not np.array_equal(single[0], s) for s in single[1:])
        if env.fake:
            assert env.calls("swag_sample_batch") >= 1
    # the first two draws of the injected run are the reference's own samples for that noise
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        model, opt = loaded()
        opt.presample(2)
        for k in range(2):
            opt.sample_parameters()
            np.testing.assert_allclose(flat(model.parameters()), g["samples"][k], rtol=RTOL, atol=ATOL)
    # a training step drops what is left of the batch and restores the training weights
    opt.presample(3)
    opt.sample_parameters()
    fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
    opt.step(fwd, bwd)
    assert opt._pre_ready == 0 and opt._pre_pending == 0 and not opt.state["__params_dirty"]


def test_deep_ensemble_announces_batches_to_swag_members(env, golden):
    """DeepEnsemble.predict tells members that can presample how many draws follow; predictions equal the
    one-by-one path."""
    g = golden("swag_steps.npz")
    x = env.t(g["xs"][0])

    def ensemble():
        pairs = []
        for m in range(2):
            model, opt = build_swag(env, g)
            for s in range(4):   # a few updates so that the deviation ring is populated
                if s == 1:
                    opt.complete_epoch()
                fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
                opt.step(fwd, bwd)
            pairs.append((model, opt))
        return bde.DeepEnsemble(pairs)

    preds = []
    for batched in (True, False):
        ens = ensemble()
        if not batched:
            for opt in ens.optimizers:
                opt.presample = lambda count: None
        noise.set_seed(77)
        with torch.no_grad():
            preds.append(ens.predict(lambda mdl: mdl(x).squeeze(-1), samples=7).cpu().numpy())
        noise.set_seed(None)
    np.testing.assert_array_equal(preds[0], preds[1])
    assert preds[0].shape[0] == 7


@pytest.mark.parametrize("batched", [True, False], ids=["presample", "one-by-one"])
def test_deep_ensemble_over_swag_members_matches_reference(env, golden, batched):
    """The reference's DeepEnsemble.predict over two SWAG members (fixture recorded from the unmodified reference
    with its noise draws on tape): the batched sampler announced by predict() and the one-by-one path both
    reproduce the reference's 4 + 3 predictions."""
    g = golden("ensemble_swag_predict.npz")
    pairs = []
    for m in range(g["inits"].shape[0]):
        model = gm.make_mlp().to(env.dev)
        gm.load_flat(model.parameters(), g["inits"][m])
        base = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
        opt = bde.SwagOptimizer(model.parameters(), base, update_interval=1, start_epoch=0, deviation_samples=4)
        for s in range(6):
            fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
            opt.step(fwd, bwd)
        if not batched:
            opt.presample = lambda count: None
        pairs.append((model, opt))
    ens = bde.DeepEnsemble(pairs)
    x = env.t(g["xs"][7])
    with noise.inject(tape(g["eps"], g["eps_sizes"])), torch.no_grad():
        preds = ens.predict(lambda mdl: mdl(x).squeeze(-1), samples=7)
    np.testing.assert_allclose(preds.cpu().numpy(), g["preds"], rtol=3e-5, atol=3e-6)
    if env.fake and batched:
        assert env.calls("swag_sample_batch") == 2 and env.calls("swag_sample") == 0


def test_swag_loads_reference_layout_checkpoint(env, golden):
    g = golden("swag_steps.npz")
    model, opt = build_swag(env, g)
    run_swag(env, model, opt, g)
    sd = copy.deepcopy(opt.state_dict())
    model2, opt2 = build_swag(env, g)
    model2.load_state_dict(model.state_dict())
    opt2.load_state_dict(sd)
    assert opt2.state["__updates"] == int(g["updates"])
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        opt2.sample_parameters()
    np.testing.assert_allclose(flat(model2.parameters()), g["samples"][0], rtol=3e-5, atol=3e-6)


# ---------------------------------------------------------------- iVON
def build_ivon(env, g):
    model = gm.make_mlp().to(env.dev)
    gm.load_flat(model.parameters(), g["init"])
    opt = bde.iVONOptimizer(model.parameters(), lr=1e-2, prior_prec=10.0, dataset_size=768, damping=1e-3,
                            mc_samples=2, augmentation=1.0, tempering=1.0)
    return model, opt


def ivon_tape(g, model):
    """The reference draws per tensor; this repo draws once per arena: regroup the tape."""
    per_call = sum(p.numel() for p in model.parameters())
    return tape(g["eps"], [per_call] * (g["eps"].size // per_call))


@pytest.mark.parametrize("capture", ["prebound", "gather", "closure-zeroes-grads"])
def test_ivon_matches_reference(env, golden, capture):
    """capture: how the gradients reach the accumulation arena — autograd accumulating straight into
    pre-bound `.grad` views, the gather-accumulate launch, or pre-binding defeated by a closure that calls
    zero_grad() itself (falls back to the gather on the same arena)."""
    g = golden("ivon_steps.npz")
    model, opt = build_ivon(env, g)
    opt.prebind_grads = capture != "gather"
    assert opt.get_base_optimizer() is opt
    with noise.inject(ivon_tape(g, model)):
        for s in range(g["losses"].size):
            fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
            if capture == "closure-zeroes-grads":
                inner = fwd

                def fwd(inner=inner):
                    model.zero_grad(set_to_none=True)
                    return inner()
            n_gathers = env.calls("mtc")
            loss = opt.step(fwd, bwd)
            if env.fake is not None:   # launches per step: 0 gathers when pre-bound, one per MC sample otherwise
                assert env.calls("mtc") - n_gathers == (0 if capture == "prebound" else opt.mc_samples)
            np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
            st = [opt.state[p] for p in model.parameters()]
            for name, key in (("mean", "means"), ("momentum", "momenta"), ("precision", "precisions")):
                got = torch.cat([x[name].reshape(-1) for x in st]).cpu().numpy()
                np.testing.assert_allclose(got, g[key][s], rtol=3e-5, atol=3e-6, err_msg=f"{name} step {s}")
            assert st[0]["delta"] is not None and st[0]["acc_grad"] is not None
        opt.sample_parameters()
        np.testing.assert_allclose(flat(model.parameters()), g["sampled"], rtol=3e-5, atol=3e-6)
    assert opt.param_groups[0]["step"] == g["losses"].size


@pytest.mark.parametrize("count,max_rows,two_groups", [(2, 99, False), (5, 99, True), (5, 2, True)])
def test_ivon_presample_equals_sequential_draws(env, golden, count, max_rows, two_groups):
    """presample(count): the batched K5 launch hands out exactly the draws — and leaves exactly the
    delta_sum — of `count` single sample_parameters() calls, with injected noise (the reference's draw order over
    calls and parameter groups) and with Philox streams, also when the buffer cap splits the request, with two
    parameter groups (round-robin stream ids), and a step() afterwards behaves as if nothing had been presampled."""
    g = golden("ivon_steps.npz")

    def trained():
        model = gm.make_mlp().to(env.dev)
        gm.load_flat(model.parameters(), g["init"])
        plist = list(model.parameters())
        params = [{"params": plist[:2]}, {"params": plist[2:], "prior_prec": 20.0}] if two_groups else plist
        opt = bde.iVONOptimizer(params, lr=1e-2, prior_prec=10.0, dataset_size=768, damping=1e-3, mc_samples=2)
        noise.set_seed(5)
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
        opt.step(fwd, bwd)    # non-trivial precision / mean
        return model, opt

    sizes = None
    gen = torch.Generator().manual_seed(11)

    def draws(batched, injected):
        nonlocal sizes
        model, opt = trained()
        if sizes is None:
            sizes = [ar["layout"].logical_size for ar in opt._arenas]
        zs = [torch.randn(sizes[i % len(sizes)], generator=torch.Generator().manual_seed(100 + i))
              for i in range(count * len(sizes))]
        it = iter(zs)
        if batched:
            opt.presample_max_bytes = 4 * sum(ar["layout"].size for ar in opt._arenas) * max_rows
            opt.presample(count)
        noise.set_seed(1234)
        out = []
        with noise.inject((lambda kind, numel: next(it)) if injected else (lambda kind, numel: None)):
            for _ in range(count):
                opt.sample_parameters()
                out.append(flat(model.parameters()).copy())
        dsum = torch.cat([ar["rows"]["delta"][:1 << 30].reshape(-1) for ar in opt._arenas]).cpu().numpy().copy()
        # a training step after prediction-time sampling: same result either way
        noise.set_seed(99)
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][1]), env.t(g["ys"][1]))
        loss = opt.step(fwd, bwd).item()
        noise.set_seed(None)
        return out, dsum, loss, flat(model.parameters()).copy()

    for injected in (True, False):
        one = draws(False, injected)
        bat = draws(True, injected)
        for a, b in zip(one[0], bat[0]):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(one[1], bat[1])
        assert one[2] == bat[2]
        np.testing.assert_array_equal(one[3], bat[3])
        assert any(not np.array_equal(one[0][0], x) for x in one[0][1:])
    if env.fake:
        assert env.calls("ivon_sample_batch") >= 1


def test_ivon_state_dict_roundtrip(env, golden):
    g = golden("ivon_steps.npz")
    model, opt = build_ivon(env, g)
    with noise.inject(ivon_tape(g, model)):
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
        opt.step(fwd, bwd)
    sd = copy.deepcopy(opt.state_dict())
    assert set(sd["state"][0].keys()) == {"mean", "momentum", "precision", "delta", "acc_grad"}
    assert sd["param_groups"][0]["step"] == 1 and sd["param_groups"][0]["lr"] == 1e-2
    model2, opt2 = build_ivon(env, g)
    opt2.load_state_dict(sd)
    for p, q in zip(model.parameters(), model2.parameters()):
        for name in ("mean", "momentum", "precision"):
            assert torch.equal(opt.state[p][name], opt2.state[q][name])
    assert opt2.state[next(iter(model2.parameters()))]["mean"].data_ptr() == opt2._arenas[0]["views"]["mean"][0].data_ptr()
    # lr schedulers act on the optimizer itself (ivorn.py:117-118)
    torch.optim.lr_scheduler.LambdaLR(opt2.get_base_optimizer(), lambda e: 0.5)
    assert opt2.param_groups[0]["lr"] == 0.5e-2


def test_ivon_philox_sampling_statistics(env):
    """Without injected noise the kernels draw Philox normals: delta ~ N(0, 1/(N prec))."""
    torch.manual_seed(0)
    model = torch.nn.Linear(256, 256).to(env.dev)
    opt = bde.iVONOptimizer(model.parameters(), lr=1e-3, prior_prec=50.0, dataset_size=1000, mc_samples=1)
    mean = flat(model.parameters()).copy()
    opt.sample_parameters()
    a = flat(model.parameters()) - mean
    opt.sample_parameters()
    b = flat(model.parameters()) - mean
    sigma = 1.0 / np.sqrt(1000 * 50.0 / 1000)
    assert abs(a.std() / sigma - 1) < 0.02 and abs(a.mean()) < 0.01 * sigma * 3
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.02  # fresh stream per call


# ---------------------------------------------------------------- BBB / Rank-1
@pytest.mark.parametrize("batched", [True, False], ids=["one-launch", "per-tensor"])
def test_bbb_rank1_matches_reference(env, golden, batched):
    """Three reference Rank-1 steps (Gaussian prior, l2_scale on the deterministic weights).  one-launch: the
    whole prior term (4 Gaussian tensors + 4 deterministic tensors) is ONE multi-tensor launch for the value
    and one for the gradients; per-tensor: a K9 / K10 launch pair per tensor."""
    g = golden("bbb_steps.npz")
    model = gm.Rank1MLP(bde.GaussianParameter).to(env.dev)
    init = {k[len("init/"):]: g[k] for k in g.files if k.startswith("init/")}
    gm.init_rank1(model, init)
    prior = bde.GaussianPrior(0.5, 0.8)
    base = torch.optim.Adam(model.parameters(), lr=1e-2)
    opt = bde.BBBOptimizer(model.parameters(), base, prior, dataset_size=100, mc_samples=2, kl_rescaling=0.5,
                           components=1, l2_scale=0.01)
    opt.batch_prior_terms = batched
    with noise.inject(tape(g["eps"], g["eps_sizes"])):
        for s in range(g["losses"].size):
            fwd, bwd = gm.mse_closures(model, env.t(g["xs"][s]), env.t(g["ys"][s]))
            loss = opt.step(fwd, bwd)
            np.testing.assert_allclose(loss.item(), g["losses"][s], rtol=1e-5)
            for name, p in model.named_parameters():
                np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"step{s}/{name}"], rtol=1e-4, atol=1e-5,
                                           err_msg=f"{name} after step {s}")
    if env.fake:
        steps = g["losses"].size
        assert env.calls("prior_terms") == (2 * steps if batched else 0)       # value + all gradients
        assert env.calls("kl_gauss") == (0 if batched else 2 * 4 * steps)     # value + grad, 4 Gaussian tensors
        assert env.calls("l2") == (0 if batched else 2 * 4 * steps)
    assert opt.sample_parameters() is None


def test_bbb_first_step_gradients_tight(env, golden):
    """One BBB step with SGD(lr=1): parameter change == gradient, at the north-star tolerance."""
    g = golden("bbb_steps.npz")
    model = gm.Rank1MLP(bde.GaussianParameter).to(env.dev)
    init = {k[len("init/"):]: g[k] for k in g.files if k.startswith("init/")}
    gm.init_rank1(model, init)
    ref = gm.Rank1MLP(bde.GaussianParameter)  # same structure, autograd-only evaluation on CPU
    gm.init_rank1(ref, init)
    eps_it = tape(g["eps"], g["eps_sizes"])
    draws = [eps_it("gauss", int(s)) for s in g["eps_sizes"][:8]]  # 2 MC samples x 4 Gaussian tensors
    x, y = torch.from_numpy(g["xs"][0]), torch.from_numpy(g["ys"][0])
    it = iter(draws)
    sp = torch.nn.functional.softplus

    def ref_layer(layer, inp):
        s = layer.s.mean + next(it) * sp(layer.s.rho)
        r = layer.r.mean + next(it) * sp(layer.r.rho)
        return layer.layer(inp * s) * r + layer.bias

    data = 0
    for _ in range(2):
        data = data + ((ref_layer(ref.l2, torch.relu(ref_layer(ref.l1, x))).squeeze(-1) - y) ** 2).mean()
    kl = 0
    for name, p in ref.named_parameters():
        if name.endswith("mean"):
            sig = sp(dict(ref.named_parameters())[name[:-4] + "rho"])
            kl = kl + (0.5 * (2 * torch.log(0.8 / sig) - 1 + (sig / 0.8) ** 2 + ((0.5 - p) / 0.8) ** 2)).sum()
        elif not name.endswith("rho"):
            kl = kl + 0.01 / 2 * p.pow(2).sum()
    loss_ref = 0.5 / 100 * kl + data / 2
    loss_ref.backward()

    base = torch.optim.SGD(model.parameters(), lr=1.0)
    opt = bde.BBBOptimizer(model.parameters(), base, bde.GaussianPrior(0.5, 0.8), dataset_size=100, mc_samples=2,
                           kl_rescaling=0.5, components=1, l2_scale=0.01)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    it2 = iter(draws)
    with noise.inject(lambda kind, numel: next(it2)):
        fwd, bwd = gm.mse_closures(model, env.t(g["xs"][0]), env.t(g["ys"][0]))
        loss = opt.step(fwd, bwd)
    np.testing.assert_allclose(loss.item(), loss_ref.item(), rtol=RTOL)
    np.testing.assert_allclose(loss.item(), g["losses"][0], rtol=RTOL)
    for (name, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        grad = (before[name] - p.detach()).cpu().numpy()
        np.testing.assert_allclose(grad, q.grad.numpy(), rtol=2e-5, atol=2e-6, err_msg=name)


def test_bbb_skips_step_on_nan_loss(env):
    model = gm.Rank1MLP(bde.GaussianParameter).to(env.dev)
    for p in model.parameters():
        torch.nn.init.constant_(p, 0.1)
    base = torch.optim.SGD(model.parameters(), lr=0.1)
    opt = bde.BBBOptimizer(model.parameters(), base, bde.GaussianPrior(0.0, 1.0), dataset_size=10)
    before = flat(model.parameters())
    called = {"bwd": 0}
    loss = opt.step(lambda: torch.tensor(float("nan"), device=env.dev),
                    lambda l: called.__setitem__("bwd", called["bwd"] + 1))
    assert loss.isnan() and called["bwd"] == 0
    np.testing.assert_array_equal(before, flat(model.parameters()))


def test_mixture_prior_kl_through_gaussian_parameter(env, golden):
    g = golden("vectors.npz")
    gp = bde.GaussianParameter(g["mix_mu"].size).to(env.dev)
    with torch.no_grad():
        gp.mean.copy_(env.t(g["mix_mu"]))
        gp.rho.copy_(env.t(g["rho"]))
    kl = gp.mean.get_parameter_kl(bde.MixturePrior(0.3, 1.0, 0.0025))
    kl.backward()
    np.testing.assert_allclose(kl.item(), g["kl_mix"], rtol=RTOL)
    np.testing.assert_allclose(gp.mean.grad.cpu().numpy(), g["kl_mix_gmu"], rtol=RTOL, atol=ATOL)


def test_gaussian_parameter_sample_and_kl_vectors(env, golden):
    g = golden("vectors.npz")
    gp = bde.GaussianParameter(g["mu"].size).to(env.dev)
    with torch.no_grad():
        gp.mean.copy_(env.t(g["mu"]))
        gp.rho.copy_(env.t(g["rho"]))
    with noise.inject(lambda kind, numel: torch.from_numpy(g["eps"])):
        w = gp.sample()
    np.testing.assert_allclose(w.detach().cpu().numpy(), g["w"], rtol=RTOL, atol=ATOL)
    w.backward(env.t(g["grad_w"]))
    np.testing.assert_array_equal(gp.mean.grad.cpu().numpy(), g["grad_mu"])
    np.testing.assert_allclose(gp.rho.grad.cpu().numpy(), g["grad_rho"], rtol=RTOL, atol=ATOL)
    gp.mean.grad = gp.rho.grad = None
    kl = gp.kl_divergence(bde.GaussianPrior(0.5, 0.8))
    (3.0 * kl).backward()  # upstream scale flows through the fused gradient kernel
    np.testing.assert_allclose(kl.item(), g["kl_gauss"], rtol=RTOL)
    np.testing.assert_allclose(gp.mean.grad.cpu().numpy(), 3.0 * g["kl_gauss_gmu"], rtol=RTOL, atol=3 * ATOL)
    np.testing.assert_allclose(gp.rho.grad.cpu().numpy(), 3.0 * g["kl_gauss_grho"], rtol=RTOL, atol=3 * ATOL)


def test_gaussian_parameter_philox_backward_regenerates_noise(env):
    """No injected noise: backward must see the same eps the forward drew (regenerated from Philox)."""
    gp = bde.GaussianParameter(4099).to(env.dev)
    gp.blundell_init()
    w = gp.sample()
    sigma = torch.nn.functional.softplus(gp.rho.detach())
    eps = (w.detach() - gp.mean.detach()) / sigma
    gw = torch.randn_like(w)
    w.backward(gw)
    expect = gw * eps * torch.sigmoid(gp.rho.detach())
    np.testing.assert_allclose(gp.rho.grad.cpu().numpy(), expect.cpu().numpy(), rtol=1e-3, atol=1e-5)
    z = eps.cpu().numpy()
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1) < 0.05


# ---------------------------------------------------------------- wrappers
def test_last_layer_wrapper_accumulates_deterministic_grads(env, golden):
    g = golden("ivon_steps.npz")
    body = torch.nn.Linear(8, 8).to(env.dev)
    head = gm.make_mlp().to(env.dev)
    ll = bde.iVONOptimizer(head.parameters(), lr=1e-2, prior_prec=10.0, dataset_size=768, mc_samples=3)
    det = torch.optim.SGD(body.parameters(), lr=0.1)
    opt = bde.LastLayerBayesianOptimizer(ll, det)
    x, y = env.t(g["xs"][0]), env.t(g["ys"][0])
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
        opt.step(fwd, bwd, grad_scaler=torch.amp.GradScaler(env.dev))
    with pytest.raises(RuntimeError):
        opt.get_base_optimizer()
    assert set(opt.state_dict().keys()) == {"ll_bayesian_optimizer", "deterministic_optimizer"}


def test_deep_ensemble_predict_matches_reference(env, golden):
    g = golden("ensemble_predict.npz")
    pairs = []
    for m in range(g["inits"].shape[0]):
        init = g["inits"][m]
        model = gm.make_mlp().to(env.dev)
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
        preds = ens.predict(lambda mdl: mdl(env.t(g["x"])).squeeze(-1), samples=7)
    np.testing.assert_allclose(preds.cpu().numpy(), g["preds"], rtol=1e-5, atol=1e-6)
    sd = ens.state_dict()
    assert set(sd.keys()) == {"models", "optimizers"} and len(sd["optimizers"]) == 2
    ens.load_state_dict(copy.deepcopy(sd))
