"""D-sharded execution: every rank holds a column slice of every particle.

Only the n*n partial squared-distance matrix crosses NVLink (one sum all-reduce of n*n
doubles per SVGD step); every other kernel on the path is local to its slice
(SURVEY.md §8e).  Rank r of R owns the columns layout.shard_bounds(D, R, r).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_dist(sc: "ops.SvgdScratch", group=None) -> None:
    """Sum the partial pair distances over the ranks holding the other column slices."""
    if world(group) > 1:
        dist.all_reduce(sc.dist, op=dist.ReduceOp.SUM, group=group)


def svgd_kernel_sharded(X: torch.Tensor, sc: "ops.SvgdScratch", l2_reg: float, kernel_grad_scale: float,
                        dataset_size: float, h_override: float = 0.0, group=None, have_partial: bool = False) -> None:
    """K1 (local partial distances) -> all-reduce of n*n fp64 -> K1b (identical on every rank): leaves
    K and A in `sc`.  With a single rank K1b runs in K1's tail (one launch).  have_partial: sc.dist
    already holds this rank's partial sums for the current X (left by the previous training-step launch)."""
    if world(group) == 1:
        if have_partial:
            ops.svgd_bandwidth(sc, l2_reg, kernel_grad_scale, dataset_size, h_override)
        else:
            ops.svgd_pairdist_bandwidth(X, sc, l2_reg, kernel_grad_scale, dataset_size, h_override)
        return
    if not have_partial:
        ops.svgd_pairdist(X, sc)
    allreduce_dist(sc, group)
    ops.svgd_bandwidth(sc, l2_reg, kernel_grad_scale, dataset_size, h_override)


def svgd_step_sharded(X: torch.Tensor, G: torch.Tensor, out: torch.Tensor, sc: "ops.SvgdScratch", l2_reg: float,
                      kernel_grad_scale: float, dataset_size: float, h_override: float = 0.0, group=None) -> torch.Tensor:
    """One SVGD posterior update on this rank's [n, D/R] slices.

    K1 (local partial distances) -> all-reduce of n*n fp64 -> K1b (identical on every rank)
    -> K2 (local).  With a single rank this is the fused two-launch path.
    """
    if world(group) == 1:
        return ops.svgd_step(X, G, out, sc, l2_reg, kernel_grad_scale, dataset_size, h_override)
    ops.svgd_pairdist(X, sc)
    allreduce_dist(sc, group)
    ops.svgd_bandwidth(sc, l2_reg, kernel_grad_scale, dataset_size, h_override)
    return ops.svgd_apply(X, G, out, sc)
