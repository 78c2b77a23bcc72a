"""MultiX container — drop-in for the reference's DeepEnsemble (src/algos/ensemble.py:8-48).

Members are independent posterior problems: each model keeps its own optimizer (any class of this
package), the container only fans state dicts and predictions out over them.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def split_samples(samples: int, members: int):
    """How many predictions each member contributes (ensemble.py:37-39): `samples // members`
    each, the FIRST member also takes the remainder."""
    share = samples // members
    return [samples - share * (members - 1)] + [share] * (members - 1)


class DeepEnsemble(nn.Module):
    def __init__(self, models_and_optimizers):
        super().__init__()
        models, optimizers = [], []
        for model, optimizer in models_and_optimizers:
            models.append(model)
            optimizers.append(optimizer)
        self.models = nn.ModuleList(models)
        self.optimizers = optimizers

    @property
    def models_and_optimizers(self):
        return list(zip(self.models, self.optimizers))

    # -- checkpoints: {"models": <ModuleList state>, "optimizers": [<optimizer state>, ...]} (ensemble.py:17-26)
    def state_dict(self, prefix='', keep_vars=False):
        out = {"models": self.models.state_dict(prefix=prefix, keep_vars=keep_vars)}
        out["optimizers"] = [optimizer.state_dict() for optimizer in self.optimizers]
        return out

    def load_state_dict(self, state_dict, strict=True):
        self.models.load_state_dict(state_dict["models"], strict=strict)
        saved = state_dict["optimizers"]
        for k in range(min(len(saved), len(self.optimizers))):   # zip semantics of the reference
            self.optimizers[k].load_state_dict(saved[k])

    def predict(self, predict_closure, samples, multisample=False):
        """`samples` predictions stacked along dim 0.  Every single prediction is preceded by that
        member's `optimizer.sample_parameters()`; a lone member that `supports_multisample` gets the
        whole request in one call when `multisample` is set (ensemble.py:28-44)."""
        lone = self.models[0] if len(self.models) == 1 else None
        if lone is not None and multisample and getattr(lone, "supports_multisample", False):
            return predict_closure(lone, n_samples=samples)

        drawn = []
        for (model, optimizer), count in zip(self.models_and_optimizers, split_samples(samples, len(self.models))):
            # optimizers that can draw a batch of posterior samples in one pass are told how many will follow
            announce = getattr(optimizer, "presample", None)
            if announce is not None:
                announce(count)
            for _ in range(count):
                optimizer.sample_parameters()
                drawn.append(predict_closure(model))
        return torch.stack(drawn)
