"""`MultiObservation` with the reference's fields (environments/wall_runner.py:11-14): proprioceptive `features` [O] (or
[B,O]) and an egocentric camera `frame` [C,H,W] (or [B,C,H,W]) of raw 0..255 pixel values (wall_runner.py:54-58)."""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class MultiObservation:
    features: torch.Tensor
    frame: torch.Tensor
