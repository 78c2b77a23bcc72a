"""D-sharded SWAG / iVON / BBB host classes (second bullet; VERDICT r1 "missing" item 7).

`SwagOptimizer`, `iVONOptimizer` and `BBBOptimizer` take `process_group=`: the ranks of the group each hold a column
slice of the weights and of every state vector.  No kernel exchanges data; what the group fixes is the noise (same
low-rank coefficients everywhere, disjoint parts of ONE Philox stream for the per-weight normals) and the value of
the BBB prior term (one scalar all-reduce).  The bar: the concatenated slices equal the unsharded run bit for bit,
whatever the rank count — on the oracle-backed ABI double over gloo (CPU) and on the CUDA library over NCCL (GPU).
A single-process variant checks the kernels' `elem0` plumbing through the classes on one GPU."""
from __future__ import annotations

import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import fake_abi
import sharded_script


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SLICED = ("swag_single", "swag_theta", "ivon_mean", "ivon_prec", "ivon_momentum", "ivon_single", "bbb_mean", "bbb_rho",
          "bbb_det", "bbb_sample")
SLICED_ROWS = ("swag_batch", "ivon_batch")


def compare(parts, full, world):
    assert parts[0]["lo"] == 0 and parts[-1]["hi"] == sharded_script.D_GLOBAL
    for r, p in enumerate(parts):
        # every rank sits where the exclusive prefix sum of the slice lengths puts it, under rank 0's key
        assert p["swag_shard"] == (p["lo"], sharded_script.D_GLOBAL, sharded_script.SEED)
        assert p["ivon_elem0"] == p["lo"] and p["bbb_offset"] == p["lo"]
        assert p["stream_position"] == full["stream_position"]
    for key in SLICED:
        got = torch.cat([p[key] for p in parts])
        assert torch.equal(got, full[key]), f"{key}: sharded x{world} differs from the unsharded run"
    for key in SLICED_ROWS:
        got = torch.cat([p[key] for p in parts], dim=1)
        assert torch.equal(got, full[key]), f"{key}: sharded x{world} differs from the unsharded run"
    # the noise of different slices is different noise (elem0 = 0 everywhere would replay rank 0's normals)
    if world > 1 and parts[0]["hi"] - parts[0]["lo"] == parts[1]["hi"] - parts[1]["lo"]:
        d0 = parts[0]["swag_single"] - parts[0]["swag_theta"]
        d1 = parts[1]["swag_single"] - parts[1]["swag_theta"]
        assert not torch.allclose(d0, d1)
    # BBB: loss = pi * (KL + L2 summed over ALL ranks) + this rank's data term (bbb.py:79-80)
    for step in range(len(full["bbb_loss"])):
        prior_full = full["bbb_loss"][step] - full["bbb_data"][step]
        for p in parts:
            np.testing.assert_allclose(p["bbb_loss"][step] - p["bbb_data"][step], prior_full, rtol=2e-5)
        np.testing.assert_allclose(sum(p["bbb_data"][step] for p in parts), full["bbb_data"][step], rtol=2e-5)


def run_case(tmp_path, world, backend, full):
    import dist_worker
    out_path = str(tmp_path / f"ew{world}")
    mp.spawn(dist_worker.run_elementwise, args=(world, free_port(), out_path, backend), nprocs=world, join=True)
    compare([torch.load(f"{out_path}.{r}") for r in range(world)], full, world)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_posteriors_gloo(tmp_path, monkeypatch, world):
    torch.set_num_threads(1)
    fake_abi.install(monkeypatch)
    full = sharded_script.run(torch.device("cpu"), 1, 0, None)
    run_case(tmp_path, world, "gloo", full)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_posteriors_nccl(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    full = sharded_script.run(torch.device("cuda", 0), 1, 0, None)
    run_case(tmp_path, world, "nccl", full)


class _FakeGroup:
    """Stands in for a process group of `world` ranks inside ONE process: the classes only reach the group through
    dist.column_shard / BBBOptimizer._place_gaussian_slices / allreduce_scalar_value, which are patched below."""

    def __init__(self, world, rank):
        self.world, self.rank = world, rank


def _single_process_parts(dev, world, monkeypatch):
    """The sharded script for rank 0 .. world-1 one after the other in this process (no torch.distributed): the
    group plumbing is replaced by its closed form, everything below it — classes, ops, C-ABI, kernels — is real."""
    from beyond_deep_ensembles_b200 import bbb as bbb_mod
    from beyond_deep_ensembles_b200 import dist as bdist
    from beyond_deep_ensembles_b200 import noise
    from beyond_deep_ensembles_b200.layout import shard_bounds

    def fake_column_shard(local_size, group=None):
        if group is None:
            return bdist.ColumnShard(bdist.SINGLE, 1, 0, 0, local_size, local_size, None)
        lo, hi = shard_bounds(sharded_script.D_GLOBAL, group.world, group.rank)
        assert hi - lo == local_size
        return bdist.ColumnShard(group, group.world, group.rank, lo, local_size, sharded_script.D_GLOBAL, sharded_script.SEED)

    def fake_place(self):
        lo, _ = shard_bounds(sharded_script.D_GLOBAL, self._group.world, self._group.rank)
        for param in self._params():
            owner = getattr(getattr(param, "get_parameter_kl", None), "__self__", None)
            if owner is not None and getattr(owner, "mean", None) is param:
                owner.column_offset, owner.noise_seed = lo, sharded_script.SEED

    monkeypatch.setattr(bdist, "column_shard", fake_column_shard)
    monkeypatch.setattr(bdist, "world", lambda group=None: group.world if isinstance(group, _FakeGroup) else 1)
    monkeypatch.setattr(bbb_mod.BBBOptimizer, "_place_gaussian_slices", fake_place)
    monkeypatch.setattr(bdist, "allreduce_scalar_value", lambda value, group: value)   # prior value: checked by the 2-rank tests
    return [sharded_script.run(dev, world, r, _FakeGroup(world, r)) for r in range(world)]


@pytest.mark.parametrize("backend", [pytest.param("cpu", id="oracle-abi"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("world", [2, 5])
def test_class_level_shard_independence_single_process(monkeypatch, backend, world):
    """Runs on ONE device: the classes place rank r's slice at elem0 = lo_r of the Philox stream, so slices generated
    one after the other equal the unsharded vectors bit for bit (SWAG single / batched draws, three iVON steps with
    Philox MC samples, Gaussian samples of BBB with their regenerated-noise backward)."""
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
    full = sharded_script.run(dev, 1, 0, None)
    parts = _single_process_parts(dev, world, monkeypatch)
    for key in SLICED:
        if key.startswith("bbb_") and key != "bbb_sample":
            continue
        assert torch.equal(torch.cat([p[key] for p in parts]), full[key]), key
    for key in SLICED_ROWS:
        assert torch.equal(torch.cat([p[key] for p in parts], dim=1), full[key]), key
    # BBB: the data-term gradients and samples are slice-local; the prior value differs (not summed here), its gradient not
    for key in ("bbb_mean", "bbb_rho", "bbb_det"):
        assert torch.equal(torch.cat([p[key] for p in parts]), full[key]), key
