"""Model closures on D-sharded parameters (SURVEY.md §8e last note, §8 f4 "sharded-closure story").

`ColumnShardedModel` all-gathers the ranks' column slices into the flat weight vector the model computes on and
reduce-scatters the gradients back, so `SVGDOptimizer` / `SwagOptimizer` / `iVONOptimizer(process_group=...)`
train a real model while every rank stores 1/R of the particles, moments and base-optimizer state.  The bar: the
job equals the plain classes on the whole model — bit for bit for the elementwise family when all ranks see the
same batch at R = 2 (averaging two equal gradients is exact), to fp32 rounding otherwise (SVGD's n x n distance
sums and the data-parallel gradient mean are added in a different order).  gloo + the oracle-backed ABI double
here, NCCL + the CUDA library on a multi-GPU box."""
from __future__ import annotations

import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import fake_abi
import sharded_closure_script as script


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


VECTORS = ("init", "swag_theta", "swag_mean", "swag_sample", "swag_pred", "ivon_mean", "ivon_prec", "svgd_init",
           "svgd_particles")


def compare(parts, full, world, split_batch):
    exact = world == 2 and not split_batch
    for key in VECTORS:
        for r, p in enumerate(parts):
            assert p[key].shape == full[key].shape, key
            if key in ("init", "svgd_init") or (exact and not key.startswith("svgd")):
                assert torch.equal(p[key], full[key]), f"{key}: rank {r} of {world} differs from the unsharded run"
            else:
                np.testing.assert_allclose(p[key].numpy(), full[key].numpy(), rtol=2e-5, atol=2e-6,
                                           err_msg=f"{key}: rank {r} of {world}")
        # every rank exported the same job-wide vectors
        assert all(torch.equal(p[key], parts[0][key]) for p in parts), key
    for key in ("swag_losses", "svgd_losses"):
        got = np.mean([p[key] for p in parts], axis=0)      # split batch: the mean of the micro-batch losses
        np.testing.assert_allclose(got, full[key], rtol=2e-5)
    # one all-gather per forward and one reduce-scatter per backward: SWAG 5 steps + 1 evaluation gather,
    # iVON 3 steps x 2 MC samples, SVGD 3 steps x 4 particles
    assert all(p["collectives"] == [2 * 5 + 1, 2 * 3 * 2, 2 * 3 * 4] for p in parts)


def run_case(tmp_path, world, backend, full, split_batch):
    import dist_worker
    out_path = str(tmp_path / f"sc{world}")
    mp.spawn(dist_worker.run_sharded_closure, args=(world, free_port(), out_path, backend, split_batch), nprocs=world,
             join=True)
    compare([torch.load(f"{out_path}.{r}") for r in range(world)], full, world, split_batch)


@pytest.mark.parametrize("world,split_batch", [(2, False), (2, True), (3, True)])
def test_sharded_closure_gloo(tmp_path, monkeypatch, world, split_batch):
    torch.set_num_threads(1)
    fake_abi.install(monkeypatch)
    full = script.run(torch.device("cpu"), 1, 0, None, split_batch)
    run_case(tmp_path, world, "gloo", full, split_batch)


def test_sharded_closure_on_a_subgroup_gloo(tmp_path, monkeypatch):
    """The column slices live on ranks 1 and 2 of a 3-rank job (process_group = a sub-group, not WORLD): group ranks,
    the broadcast source and every collective must be the sub-group's."""
    import dist_worker
    torch.set_num_threads(1)
    fake_abi.install(monkeypatch)
    full = script.run(torch.device("cpu"), 1, 0, None, True)
    out_path = str(tmp_path / "sub")
    mp.spawn(dist_worker.run_sharded_closure, args=(3, free_port(), out_path, "gloo", True, True), nprocs=3, join=True)
    assert torch.load(f"{out_path}.0") is None
    compare([torch.load(f"{out_path}.{r}") for r in (1, 2)], full, 2, True)


@pytest.mark.gpu
@pytest.mark.parametrize("world,split_batch", [(2, False), (2, True), (4, True)])
def test_sharded_closure_nccl(tmp_path, world, split_batch):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    full = script.run(torch.device("cuda", 0), 1, 0, None, split_batch)
    run_case(tmp_path, world, "nccl", full, split_batch)


@pytest.mark.parametrize("backend", [pytest.param("cpu", id="oracle-abi"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def test_one_rank_group_equals_the_plain_classes(monkeypatch, backend):
    """ONE device, no torch.distributed: SWAG / iVON / SVGD over ColumnShardedModel's flat parameter and closures
    (world = 1) against the plain classes over the model's own tensors — same arena layout, same Philox counters,
    so the elementwise family agrees bit for bit and SVGD to rounding (one [n, 640] tensor instead of four)."""
    if backend == "cuda":
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
        from beyond_deep_ensembles_b200 import _lib
        _lib.get()
        dev = torch.device("cuda", 0)
    else:
        torch.set_num_threads(1)
        fake_abi.install(monkeypatch)
        dev = torch.device("cpu")
    full = script.run(dev, 1, 0, None, False)
    wrapped = script.run(dev, 1, 0, None, False, wrap_single=True)
    for key in VECTORS:
        if key in ("init", "svgd_init") or (backend == "cpu" and not key.startswith("svgd")):
            assert torch.equal(wrapped[key], full[key]), key
        else:   # fp32 rtol 1e-5 / atol 1e-6 (north star); on the GPU also for the elementwise family (library GEMMs in the closures)
            np.testing.assert_allclose(wrapped[key].numpy(), full[key].numpy(), rtol=1e-5, atol=1e-6, err_msg=key)
    np.testing.assert_allclose(wrapped["swag_losses"], full["swag_losses"], rtol=1e-6)
    np.testing.assert_allclose(wrapped["svgd_losses"], full["svgd_losses"], rtol=1e-5)
    assert wrapped["collectives"] == [0, 0, 0]


def test_single_rank_closure_is_the_plain_model(monkeypatch):
    """world = 1 (process_group=None): no collective; the closures still run the model on `full`, hand the gradient
    to `param.grad` with autograd's accumulate semantics, and survive a closure that replaces the .grad tensors."""
    import beyond_deep_ensembles_b200 as bde
    import golden_models as gm
    fake_abi.install(monkeypatch)
    torch.manual_seed(0)
    model, twin = gm.make_mlp(), gm.make_mlp()
    twin.load_state_dict(model.state_dict())
    sm = bde.ColumnShardedModel(model)
    assert sm.world == 1 and sm.shard == sm.layout.size == sm.param.numel() and sm.shard % 64 == 0
    x, y = torch.randn(16, 8), torch.randn(16)
    fwd, bwd = sm.closures(*gm.mse_closures(model, x, y))
    tf, tb = gm.mse_closures(twin, x, y)
    tb(tf())
    want = sm.layout.from_logical(torch.cat([p.grad.reshape(-1) for p in twin.parameters()]))
    bwd(fwd())
    assert torch.equal(sm.param.grad, want)
    bwd(fwd())                                               # accumulates like autograd
    assert torch.equal(sm.param.grad, 2 * want)
    # a closure that drops the pre-bound .grad tensors (zero_grad(set_to_none=True) inside it)
    f0, b0 = gm.mse_closures(model, x, y)

    def bwd_replacing(loss):
        for p in model.parameters():
            p.grad = None
        b0(loss)
    fwd2, bwd2 = sm.closures(f0, bwd_replacing)
    sm.param.grad = None
    bwd2(fwd2())
    assert torch.equal(sm.param.grad, want)
    # an update of the slice reaches the model at the next forward, padding columns stay zero
    with torch.no_grad():
        sm.param.add_(1.0)
    fwd()
    for p, q in zip(model.parameters(), twin.parameters()):
        assert torch.equal(p.detach(), q.detach() + 1.0)
    assert sm.collectives == 0
    with pytest.raises(TypeError):
        bde.ColumnShardedModel(gm.make_mlp().double())


def test_sharded_optimizers_refuse_an_active_grad_scaler(monkeypatch):
    """A D-sharded optimizer sees one rank's columns: the scaler's non-finite check would be taken per rank and the
    ranks could disagree about skipping a step — step() raises instead (single-rank jobs keep their AMP path)."""
    import beyond_deep_ensembles_b200 as bde
    from beyond_deep_ensembles_b200 import dist as bdist
    fake_abi.install(monkeypatch)

    class Scaler:
        def is_enabled(self):
            return True

    w = torch.nn.Parameter(torch.randn(64))
    fwd, bwd = (lambda: (w ** 2).sum()), (lambda loss: loss.backward())
    swag = bde.SwagOptimizer([w], torch.optim.SGD([w], lr=0.1), update_interval=1, deviation_samples=2)
    ivon = bde.iVONOptimizer([torch.nn.Parameter(torch.randn(64))], lr=0.01, prior_prec=1.0, dataset_size=10, mc_samples=1)
    swag._shard.world = 2
    ivon._arenas[0]["shard"].world = 2
    for opt in (swag, ivon):
        with pytest.raises(ValueError, match="GradScaler"):
            opt.step(fwd, bwd, grad_scaler=Scaler())
    monkeypatch.setattr(bdist, "world", lambda group=None: 2)
    v = torch.nn.Parameter(torch.randn(64))
    svgd = bde.SVGDOptimizer.__new__(bde.SVGDOptimizer)        # the guard sits in front of everything step() does
    torch.optim.Optimizer.__init__(svgd, [v], {})
    svgd.state["__particle_count"], svgd.state["__base_optimizer"] = 2, torch.optim.SGD([v], lr=0.1)
    svgd._group = object()
    from beyond_deep_ensembles_b200.layout import ParamLayout
    svgd._layout = ParamLayout([v])
    with pytest.raises(ValueError, match="GradScaler"):
        svgd.step(fwd, bwd, grad_scaler=Scaler())
