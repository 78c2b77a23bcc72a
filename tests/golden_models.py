"""Small self-contained models shared by oracle/gen_golden.py (which drives the REFERENCE
optimizers on them) and the parity tests (which drive this repo's optimizers on them).

Nothing here imports the reference: the GPU box has no /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def make_mlp(in_dim: int = 8, hidden: int = 50) -> nn.Sequential:
    """The UCI regression body of the reference (experiments/uci/models.py:121-127), D = 501."""
    return nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, 1))


def flat_params(params) -> np.ndarray:
    return torch.cat([p.detach().reshape(-1).cpu() for p in params]).numpy()


def load_flat(params, flat: np.ndarray) -> None:
    """Copy a flat vector into the parameters (in place, device preserved)."""
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(torch.from_numpy(np.ascontiguousarray(flat[off:off + n])).view_as(p).to(p.device))
            off += n
    assert off == flat.size


class Rank1Linear(nn.Module):
    """Rank-1 VI linear layer with the structure of the reference's (src/algos/rank1.py:9-64),
    parameterised by the GaussianParameter class so that the same definition runs on the
    reference's class (golden generation) and on this repo's class (parity test)."""

    def __init__(self, gp_cls, in_features: int, out_features: int):
        super().__init__()
        self.layer = nn.Linear(in_features, out_features, bias=False)
        self.s = gp_cls(in_features)
        self.r = gp_cls(out_features)
        self.bias = nn.Parameter(torch.zeros(out_features))

    def forward(self, x):
        s = self.s.sample()
        r = self.r.sample()
        return self.layer(x * s) * r + self.bias


class Rank1MLP(nn.Module):
    def __init__(self, gp_cls, in_dim: int = 8, hidden: int = 16):
        super().__init__()
        self.l1 = Rank1Linear(gp_cls, in_dim, hidden)
        self.l2 = Rank1Linear(gp_cls, hidden, 1)

    def forward(self, x):
        return self.l2(torch.relu(self.l1(x)))


def init_rank1(model: Rank1MLP, flat_by_name: dict) -> None:
    """Deterministic init from a dict name -> array (state_dict order is not relied upon)."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(torch.from_numpy(flat_by_name[name]).view_as(p).to(p.device))


def mse_closures(model, x, y):
    """(forward_closure, backward_closure) in the reference's convention (algo.py:19-29)."""

    def forward():
        return ((model(x).squeeze(-1) - y) ** 2).mean()

    def backward(loss):
        loss.backward()

    return forward, backward
