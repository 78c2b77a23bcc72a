"""MultiX container — drop-in for the reference's DeepEnsemble (src/algos/ensemble.py:8-48)."""
from __future__ import annotations

import torch
import torch.nn as nn


class DeepEnsemble(nn.Module):
    """Models together with their optimizers; members are independent posterior problems."""

    def __init__(self, models_and_optimizers):
        super().__init__()
        pairs = list(models_and_optimizers)
        self.models = nn.ModuleList([m for m, _ in pairs])
        self.optimizers = [o for _, o in pairs]

    def state_dict(self, prefix='', keep_vars=False):
        return {
            "models": self.models.state_dict(prefix=prefix, keep_vars=keep_vars),
            "optimizers": [o.state_dict() for o in self.optimizers],
        }

    def load_state_dict(self, state_dict, strict=True):
        self.models.load_state_dict(state_dict["models"], strict=strict)
        for optimizer, optimizer_state in zip(self.optimizers, state_dict["optimizers"]):
            optimizer.load_state_dict(optimizer_state)

    def predict(self, predict_closure, samples, multisample=False):
        """`samples` predictions: samples // members per member, the first member takes the
        remainder; every prediction is preceded by optimizer.sample_parameters() (ensemble.py:28-44)."""
        if len(self.models) == 1 and getattr(self.models[0], "supports_multisample", False) and multisample:
            return predict_closure(self.models[0], n_samples=samples)

        members = len(self.models)
        per_model = samples // members
        output = []
        for i, (model, optimizer) in enumerate(self.models_and_optimizers):
            count = per_model if i > 0 else samples - (members - 1) * per_model
            for _ in range(count):
                optimizer.sample_parameters()
                output.append(predict_closure(model))
        return torch.stack(output)

    @property
    def models_and_optimizers(self):
        return list(zip(self.models, self.optimizers))
