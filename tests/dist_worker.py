"""Worker of tests/test_sharding_gloo.py: one process per rank, gloo backend, the package's own
sharding code (dist.svgd_step_sharded) over the oracle-backed ABI double."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def run_timeout(rank: int, world: int, port: int, out_path: str):
    """A rank that never reaches its exchange: the waiting rank's kernel gives up after BDE_PEER_TIMEOUT_S, leaves
    K / A untouched, marks the workspace failed (later exchanges return at once) and PeerSet.check() raises."""
    import time
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), BDE_PEER_TIMEOUT_S="2")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from beyond_deep_ensembles_b200 import _lib
    from beyond_deep_ensembles_b200 import dist as bdist
    from beyond_deep_ensembles_b200 import ops
    n, D = 10, 300_000
    g = torch.Generator().manual_seed(7 + rank)
    X = (torch.randn(n, D, generator=g) * 0.05).to(dev)
    sc = ops.SvgdScratch.allocate(n, dev)
    assert bdist.enable_peer_exchange(sc)
    ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)      # a good exchange on every rank
    torch.cuda.synchronize()
    sc.peers.check()
    K_good = sc.K.clone()
    res = {"rank": rank}
    dist.barrier()
    if rank == 0:
        t0 = time.time()
        ops.svgd_pairdist_bandwidth(X * 1.5, sc, 0.01, 1.0, 50000.0)   # the peer never launches this one
        torch.cuda.synchronize()
        res["waited_s"] = time.time() - t0
        res["K_unchanged"] = bool(torch.equal(sc.K, K_good))
        res["dist_poisoned"] = bool(torch.isnan(sc.dist).any())
        try:
            sc.peers.check()
            res["raised"] = False
        except _lib.BdeError as e:
            res["raised"] = True
            res["message"] = str(e)
        t0 = time.time()
        ops.svgd_pairdist_bandwidth(X, sc, 0.01, 1.0, 50000.0)          # failed workspace: no second wait
        torch.cuda.synchronize()
        res["second_wait_s"] = time.time() - t0
        res["status"] = sc.peers.status()
    else:
        time.sleep(5.0)
    torch.save(res, f"{out_path}.{rank}")
    dist.barrier()
    bdist.shutdown_peer_exchange()
    dist.destroy_process_group()


def run_rank_local(rank: int, world: int, port: int, out_path: str, backend: str = "nccl"):
    """SVGDOptimizer(process_group=None) under an initialised default group: no collective, rank-local bandwidth."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        import fake_abi
        fake_abi.install(_Patch())
    import beyond_deep_ensembles_b200 as bde
    import golden_models as gm

    def one_step():
        torch.manual_seed(100 + rank)           # DDP-style: every rank its own particles / data
        model = gm.make_mlp().to(dev)
        base = torch.optim.SGD(model.parameters(), lr=1e-2)

        def reset():
            for m in model:
                if hasattr(m, "reset_parameters"):
                    m.reset_parameters()
        opt = bde.SVGDOptimizer(model.parameters(), reset, base, particle_count=5, dataset_size=768, l2_reg=0.01)
        x, y = torch.randn(32, 8, device=dev), torch.randn(32, device=dev)
        fwd, bwd = gm.mse_closures(model, x, y)
        opt.step(fwd, bwd)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        return opt, float(opt._scratch.info[0])

    calls = []
    orig = dist.all_reduce

    def spy(*a, **k):
        calls.append("all_reduce")
        return orig(*a, **k)
    dist.all_reduce = spy
    opt, h_ddp = one_step()
    dist.all_reduce = orig
    res = {"h_ddp": h_ddp, "peers": opt._scratch.peers, "abi_collectives": len(calls)}
    dist.barrier()
    dist.destroy_process_group()
    _, res["h_alone"] = one_step()              # the same seeds without torch.distributed
    torch.save(res, f"{out_path}.{rank}")


def run(rank: int, world: int, port: int, n: int, D: int, out_path: str, backend: str = "gloo", peer: bool = False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        fake = None
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        import fake_abi
        fake = fake_abi.install(_Patch())
    from beyond_deep_ensembles_b200 import dist as bdist
    from beyond_deep_ensembles_b200 import ops
    from beyond_deep_ensembles_b200.layout import shard_bounds

    g = torch.Generator().manual_seed(1234)  # every rank builds the same global problem ...
    X = torch.randn(n, D, generator=g) * (0.05 * (1 + 0.1 * torch.arange(n).float())).unsqueeze(1)
    G = 1e-3 * torch.randn(n, D, generator=g)
    lo, hi = shard_bounds(D, world, rank)          # ... and keeps only its column slice
    Xl, Gl = X[:, lo:hi].contiguous().to(dev), G[:, lo:hi].contiguous().to(dev)
    out = torch.empty_like(Xl)
    sc = ops.SvgdScratch.allocate(n, dev)
    extra = {}
    if peer:
        # in-kernel exchange over peer memory: the sharded step must be the single-GPU launch sequence
        assert bdist.enable_peer_exchange(sc), "CUDA IPC peer mapping failed"
        from beyond_deep_ensembles_b200 import _lib
        for _ in range(3):   # epochs advance; parity double-buffering is exercised
            l0 = _lib.launch_count
            bdist.svgd_step_sharded(Xl, Gl, out, sc, 0.01, 1.0, 50000.0)
            extra["abi_calls_per_step"] = _lib.launch_count - l0
        # a second scratch on the same peer set (MultiX members share the exchange buffers)
        sc2 = ops.SvgdScratch.allocate(n, dev)
        assert bdist.enable_peer_exchange(sc2)
        ops.svgd_pairdist(Xl, sc2)
        extra["dist2"] = sc2.dist.cpu()
        # training-step kernels: K2 + SGD + next pair distances + cross-rank sum + K1b in ONE launch per step
        Xt, buf = Xl.clone(), torch.zeros(hi - lo, device=dev)
        nk = ops.NextKernel(True, 0.01, 1.0, 50000.0)
        kw = dict(lr=0.05, momentum=0.9, nesterov=True, weight_decay=3e-4)
        if 2 <= n <= ops.NEXT_KERNEL_MAX_PARTICLES:
            ops.svgd_pairdist_bandwidth(Xt, sc2, 0.01, 1.0, 50000.0)
            for s in range(3):
                ops.svgd_apply_sgd(Xt, Gl, sc2, buf, buf_initialized=(s > 0), next_kernel=nk, **kw)
            extra["train_X"], extra["train_dist"], extra["train_K"] = Xt.cpu(), sc2.dist.cpu(), sc2.K.cpu()
        torch.cuda.synchronize()
        extra["peer_status"] = sc.peers.status()
    else:
        bdist.svgd_step_sharded(Xl, Gl, out, sc, 0.01, 1.0, 50000.0)
    torch.save({"lo": lo, "hi": hi, "out": out.cpu(), "dist": sc.dist.cpu(), "sel": sc.sel.cpu(), "K": sc.K.cpu(),
                "calls": list(fake.calls) if fake else None, **extra}, f"{out_path}.{rank}")
    dist.barrier()
    bdist.shutdown_peer_exchange()
    dist.destroy_process_group()


def run_elementwise(rank: int, world: int, port: int, out_path: str, backend: str = "gloo"):
    """D-sharded SWAG / iVON / BBB host classes (process_group=...): every rank runs tests/sharded_script.py on its
    column slice; the test compares the concatenated slices with the unsharded run."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        import fake_abi
        fake_abi.install(_Patch())
    import sharded_script
    from beyond_deep_ensembles_b200 import dist as bdist
    res = sharded_script.run(dev, world, rank, dist.group.WORLD)
    # the SPMD contract held: every rank stands at the same Philox stream position
    bdist.check_noise_in_step(bdist.column_shard(64, dist.group.WORLD))
    if dev.type == "cuda":
        torch.cuda.synchronize()
    torch.save(res, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def run_sharded_closure(rank: int, world: int, port: int, out_path: str, backend: str = "gloo", split_batch: bool = True,
                        subgroup: bool = False):
    """D-sharded optimizers over ColumnShardedModel closures (all-gather weights / reduce-scatter gradients): every
    rank runs tests/sharded_closure_script.py; the test compares with the plain classes on the whole model.
    subgroup: the column slices live on ranks 1 .. world-1 only (a sub-group of the job; rank 0 stays out)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        import fake_abi
        fake_abi.install(_Patch())
    import sharded_closure_script
    from beyond_deep_ensembles_b200 import dist as bdist
    if subgroup:
        grp = dist.new_group(list(range(1, world)))      # collective over the whole job
        res = sharded_closure_script.run(dev, world - 1, rank - 1, grp, split_batch) if rank > 0 else None
    else:
        res = sharded_closure_script.run(dev, world, rank, dist.group.WORLD, split_batch)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    torch.save(res, f"{out_path}.{rank}")
    dist.barrier()
    bdist.shutdown_peer_exchange()
    dist.destroy_process_group()
