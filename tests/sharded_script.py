"""One script, run twice: by every rank of a D-sharded job on its column slice (tests/dist_worker.py) and by the
test process on the whole vector (world = 1).  SURVEY.md §8e, second bullet: the elementwise family needs no
collective, so the concatenated slices must equal the unsharded run BIT FOR BIT — the low-rank coefficients of SWAG
are the same on every rank, the per-weight normals are the ranks' disjoint parts of one Philox stream, and the BBB
prior term's value is summed over the ranks.  Noise is Philox inside the kernels (no injection)."""
from __future__ import annotations

import torch

D_GLOBAL = 64 * 37      # slices of shard_bounds() are whole 64-element blocks: concatenated rank arenas = the global arena
SEED = 4242


def run(dev, world: int, rank: int, group):
    import beyond_deep_ensembles_b200 as bde
    from beyond_deep_ensembles_b200 import noise
    from beyond_deep_ensembles_b200.layout import shard_bounds

    noise.set_seed(SEED + (0 if group is None else 1000 * rank))   # D-sharded: ranks seeded differently ON PURPOSE
    lo, hi = shard_bounds(D_GLOBAL, world, rank)
    g = torch.Generator().manual_seed(7)
    W0 = torch.randn(D_GLOBAL, generator=g)
    Cv = torch.randn(D_GLOBAL, generator=g)
    c = Cv[lo:hi].to(dev)
    out = {"lo": lo, "hi": hi}

    def closures(param_fn):
        def fwd():
            return (0.5 * (param_fn() * c) ** 2).sum()     # separable: a slice's gradient needs the slice only

        def bwd(loss):
            loss.backward()
        return fwd, bwd

    # ---------------------------------------------------------------- SWAG (swag.py:53-58, 91-114)
    w = torch.nn.Parameter(W0[lo:hi].clone().to(dev))
    base = torch.optim.SGD([w], lr=0.1)
    opt = bde.SwagOptimizer([w], base, update_interval=1, deviation_samples=4, process_group=group)
    out["swag_shard"] = (opt._shard.elem0, opt._shard.total, opt._shard.seed)
    fwd, bwd = closures(lambda: w)
    for _ in range(6):
        opt.step(fwd, bwd)
    opt.sample_parameters()
    out["swag_single"] = w.detach().cpu().clone()
    opt.presample(3)
    draws = []
    for _ in range(3):
        opt.sample_parameters()
        draws.append(w.detach().cpu().clone())
    out["swag_batch"] = torch.stack(draws)
    opt.step(fwd, bwd)                                      # restores the training weights
    out["swag_theta"] = w.detach().cpu().clone()

    # ---------------------------------------------------------------- iVON (ivorn.py:66-115)
    v = torch.nn.Parameter(W0[lo:hi].clone().to(dev))
    opt = bde.iVONOptimizer([v], lr=0.01, prior_prec=10.0, dataset_size=1000, mc_samples=2, damping=1e-3,
                            process_group=group)
    fwd, bwd = closures(lambda: v)
    for _ in range(3):
        opt.step(fwd, bwd)
    st = opt.state[v]
    out["ivon_mean"], out["ivon_prec"] = st["mean"].detach().cpu().clone(), st["precision"].detach().cpu().clone()
    out["ivon_momentum"] = st["momentum"].detach().cpu().clone()
    opt.sample_parameters()
    out["ivon_single"] = v.detach().cpu().clone()
    opt.presample(3)
    draws = []
    for _ in range(3):
        opt.sample_parameters()
        draws.append(v.detach().cpu().clone())
    out["ivon_batch"] = torch.stack(draws)
    out["ivon_elem0"] = opt._arenas[0]["shard"].elem0

    # ---------------------------------------------------------------- BBB (util.py:170-183, bbb.py:69-80)
    gp = bde.GaussianParameter((hi - lo,), device=dev)
    det = torch.nn.Parameter(Cv[lo:hi].clone().to(dev))
    with torch.no_grad():
        gp.mean.copy_(0.1 * W0[lo:hi].to(dev))
        gp.rho.copy_(-3.0 + 0.2 * Cv[lo:hi].to(dev))
    params = [gp.mean, gp.rho, det]
    base = torch.optim.SGD(params, lr=0.05)
    opt = bde.BBBOptimizer(params, base, prior=bde.GaussianPrior(0.0, 1.0), dataset_size=100, l2_scale=0.01,
                           process_group=group)
    out["bbb_offset"] = int(gp.column_offset)
    data = {}

    def fwd_bbb():
        data["loss"] = ((gp.sample() + det) * c).pow(2).sum()
        return data["loss"]

    losses, datas = [], []
    for _ in range(2):
        loss = opt.step(fwd_bbb, bwd)
        losses.append(float(loss.detach()))
        datas.append(float(data["loss"].detach()))
    out["bbb_loss"], out["bbb_data"] = losses, datas
    out["bbb_mean"], out["bbb_rho"] = gp.mean.detach().cpu().clone(), gp.rho.detach().cpu().clone()
    out["bbb_det"] = det.detach().cpu().clone()
    out["bbb_sample"] = gp.sample().detach().cpu().clone()
    out["stream_position"] = noise.stream_position()
    return out
