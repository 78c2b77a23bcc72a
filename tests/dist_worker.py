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
