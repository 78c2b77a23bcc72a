"""Stand-in models for the whole-step measurements (tools/whole_step.py).

The reference's architectures (src/architectures) are out of scope and do not travel to the GPU
box; the optimizer path only needs closures whose parameter list has the right shape.  These are
independent definitions with the same parameter counts as the BASELINE.json configs:

  C1  UCI MLP                 D = 501        (tests/golden_models.make_mlp)
  C2  CIFAR ResNet-20 + FRN   D = 273,610    (96 tensors)
  C3  ResNet-50, fc -> 182    D = 23,880,950 (torchvision, random init)
  C4  DistilBERT + head       D = 66,955,010 (transformers, random init)
"""
from __future__ import annotations

import torch
import torch.nn as nn


class FRN(nn.Module):
    """Filter response normalisation with a thresholded linear unit (3 parameters per channel)."""

    def __init__(self, channels: int, eps: float = 1e-6):
        super().__init__()
        shape = (1, channels, 1, 1)
        self.gamma = nn.Parameter(torch.ones(shape))
        self.beta = nn.Parameter(torch.zeros(shape))
        self.tau = nn.Parameter(torch.zeros(shape))
        self.eps = eps

    def forward(self, x):
        nu2 = x.square().mean(dim=(2, 3), keepdim=True)
        return torch.maximum(self.gamma * x * torch.rsqrt(nu2 + self.eps) + self.beta, self.tau)


class _Block(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1)
        self.n1 = FRN(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.n2 = FRN(cout)
        self.skip = nn.Conv2d(cin, cout, 1, stride, 0, bias=False) if stride != 1 else None
        self.act = nn.SiLU()

    def forward(self, x):
        y = self.n2(self.c2(self.act(self.n1(self.c1(x)))))
        return self.act(y + (x if self.skip is None else self.skip(x)))


class ResNet20FRN(nn.Module):
    """CIFAR-style ResNet-20 (3 stages x 3 basic blocks, 16/32/64 channels), FRN + swish."""

    def __init__(self, classes: int = 10):
        super().__init__()
        layers = [nn.Conv2d(3, 16, 3, 1, 1)]
        cin = 16
        for cout, stride in ((16, 1), (32, 2), (64, 2)):
            for b in range(3):
                layers.append(_Block(cin, cout, stride if b == 0 else 1))
                cin = cout
        self.body = nn.Sequential(*layers)
        self.head = nn.Linear(64, classes)

    def forward(self, x):
        return self.head(self.body(x).mean(dim=(2, 3)))


def resnet50_fc182():
    import torchvision
    m = torchvision.models.resnet50(weights=None)
    m.fc = nn.Linear(m.fc.in_features, 182)
    return m


class DistilBertClassifier(nn.Module):
    """Random-init DistilBERT body + [768 -> 768 -> 2] head (the shape of the reference's bert.py)."""

    def __init__(self, head: nn.Module | None = None):
        super().__init__()
        from transformers import DistilBertConfig, DistilBertModel
        self.body = DistilBertModel(DistilBertConfig())
        self.head = head if head is not None else nn.Sequential(nn.Linear(768, 768), nn.ReLU(), nn.Linear(768, 2))

    def forward(self, ids, mask):
        h = self.body(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0]
        return self.head(h)


def count(model: nn.Module) -> tuple[int, int]:
    ps = list(model.parameters())
    return sum(p.numel() for p in ps), len(ps)
