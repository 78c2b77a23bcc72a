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


def run(rank: int, world: int, port: int, n: int, D: int, out_path: str, backend: str = "gloo"):
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
    bdist.svgd_step_sharded(Xl, Gl, out, sc, 0.01, 1.0, 50000.0)
    torch.save({"lo": lo, "hi": hi, "out": out.cpu(), "dist": sc.dist.cpu(), "sel": sc.sel.cpu(), "K": sc.K.cpu(),
                "calls": list(fake.calls) if fake else None}, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()
