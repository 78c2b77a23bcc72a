"""One script, run twice (like tests/sharded_script.py): by every rank of a D-sharded job through
`ColumnShardedModel` closures — the model runs on all-gathered weights, the gradients are reduce-scattered onto the
rank's column slice, the D-sharded optimizer updates that slice — and by the test process with the plain classes
on the whole model (world = 1, group None).  SURVEY.md §8e last note / §8 f4 (sharded closure).

`split_batch`: every rank takes its own equal part of the batch (data parallelism over the same group; the
averaged gradients equal the full-batch gradient to rounding); otherwise all ranks see the whole batch (the
averaged gradients are the full-batch gradient exactly at world = 2)."""
from __future__ import annotations

import torch

import golden_models as gm

SEED = 991
BATCH = 48


def _reset_fn(model):
    def reset():
        for m in model:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()
    return reset


def _data(dev, world, rank, split_batch):
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(BATCH, 8, generator=g), torch.randn(BATCH, generator=g)
    if split_batch and world > 1:
        per = BATCH // world
        x, y = x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]
    return x.to(dev), y.to(dev)


def run(dev, world: int, rank: int, group, split_batch: bool, wrap_single: bool = False):
    """Returns full-length vectors on every rank (the sharded run all-gathers its slices for the comparison).
    wrap_single (with group None): the optimizers still run over ColumnShardedModel's one flat parameter and its
    closures, as a one-rank "group" — no collective, everything else as in a sharded job."""
    import beyond_deep_ensembles_b200 as bde
    from beyond_deep_ensembles_b200 import noise

    noise.set_seed(SEED + (0 if group is None else 1000 * rank))   # ranks seeded differently on purpose
    x, y = _data(dev, world, rank, split_batch)
    out = {}

    def build():
        torch.manual_seed(3 + (0 if group is None else 17 * rank))  # rank 0's init is what the job uses (broadcast)
        model = gm.make_mlp().to(dev)
        if group is None and not wrap_single:
            return model, None, list(model.parameters())
        sm = bde.ColumnShardedModel(model, group)
        return model, sm, [sm.param]

    def closures(model, sm):
        fwd, bwd = gm.mse_closures(model, x, y)
        return (fwd, bwd) if sm is None else sm.closures(fwd, bwd)

    def export(sm, opt_layout, rows):
        """[k, arena] state rows of this run -> [k, D] in parameters_to_vector order."""
        rows = rows.reshape(-1, rows.shape[-1])
        if sm is None:
            return opt_layout.to_logical(rows).cpu()
        return sm.logical(sm.gather_rows(rows)).cpu()

    # ---------------------------------------------------------------- SWAG over a sharded SGD (ZeRO-1 arrangement)
    model, sm, params = build()
    out["init"] = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().clone()
    base = torch.optim.SGD(params, lr=0.05, momentum=0.9)
    opt = bde.SwagOptimizer(params, base, update_interval=1, deviation_samples=3, process_group=group)
    fwd, bwd = closures(model, sm)
    losses = [float(opt.step(fwd, bwd).detach()) for _ in range(5)]
    out["swag_losses"] = losses
    out["swag_theta"] = export(sm, opt._layout, opt._theta)
    out["swag_mean"] = export(sm, opt._layout, opt._mean)
    opt.sample_parameters()
    if sm is not None:
        sm.gather()                                         # evaluation forward on the drawn weights
    with torch.no_grad():
        xe, _ = _data(dev, 1, 0, False)
        out["swag_pred"] = model(xe).squeeze(-1).cpu().clone()
    out["swag_sample"] = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().clone()
    counts = [0 if sm is None else sm.collectives]

    # ---------------------------------------------------------------- iVON, 2 MC samples per step (accumulated gradients)
    model, sm, params = build()
    opt = bde.iVONOptimizer(params, lr=0.01, prior_prec=10.0, dataset_size=1000, mc_samples=2, damping=1e-3,
                            process_group=group)
    fwd, bwd = closures(model, sm)
    for _ in range(3):
        opt.step(fwd, bwd)
    ar = opt._arenas[0]
    out["ivon_mean"] = export(sm, ar["layout"], ar["rows"]["mean"])
    out["ivon_prec"] = export(sm, ar["layout"], ar["rows"]["precision"])
    counts.append(0 if sm is None else sm.collectives)

    # ---------------------------------------------------------------- SVGD, 4 particles, shared SGD (fused into K2f)
    model, sm, params = build()
    base = torch.optim.SGD(params, lr=0.02, momentum=0.9, nesterov=True, weight_decay=1e-4)
    reset = _reset_fn(model) if sm is None else sm.reset_closure(_reset_fn(model))
    torch.manual_seed(11 + (0 if group is None else 5 * rank))
    opt = bde.SVGDOptimizer(params, reset, base, particle_count=4, dataset_size=BATCH, l2_reg=0.01, process_group=group)
    fwd, bwd = closures(model, sm)
    out["svgd_init"] = export(sm, opt._layout, opt._X)
    before = 0 if sm is None else sm.collectives
    out["svgd_losses"] = [float(opt.step(fwd, bwd).detach()) for _ in range(3)]
    out["svgd_particles"] = export(sm, opt._layout, opt._X)
    counts.append(0 if sm is None else sm.collectives - before)
    out["collectives"] = counts     # all-gathers + reduce-scatters of the three sections
    return out
