"""Parameter containers for the engine-backed MLPs (reference: networks/core.py:6-10 `mlp`)."""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class ArenaLinear(nn.Module):
    """Stands in for `nn.Linear(in_features, out_features)`: same parameter names (`weight` [out,in], `bias` [out]),
    same default initialisation U(-1/sqrt(in), 1/sqrt(in)) -- but the two parameters may be VIEWS into a flat HBM
    arena that the CUDA engine reads and updates in place."""

    def __init__(self, in_features: int, out_features: int, weight: torch.Tensor = None, bias: torch.Tensor = None,
                 device=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        if weight is None:
            weight = torch.empty(out_features, in_features, device=device)
            bias = torch.empty(out_features, device=device)
        assert tuple(weight.shape) == (out_features, in_features) and tuple(bias.shape) == (out_features,)
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)
        self.reset_parameters()

    def reset_parameters(self):
        k = 1.0 / math.sqrt(self.in_features)
        with torch.no_grad():
            self.weight.uniform_(-k, k)
            self.bias.uniform_(-k, k)

    def forward(self, x):   # plain torch path, for inspection only; the engine never calls it
        return F.linear(x, self.weight, self.bias)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, arena-backed"


def mlp(neurons: List[int], device=None) -> nn.ModuleList:
    """ModuleList of linear layers [n0->n1, n1->n2, ...] (reference networks/core.py:6-10)."""
    return nn.ModuleList([ArenaLinear(i, o, device=device) for i, o in zip(neurons[:-1], neurons[1:])])
