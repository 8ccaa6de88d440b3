"""Batch container on the sampler/score-model boundary.

Mirrors ``fdiff.utils.dataclasses.DiffusableBatch`` (src/fdiff/utils/dataclasses.py:7-18).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class DiffusableBatch:
    X: torch.Tensor  # [B, max_len, n_channels]
    y: Optional[torch.Tensor] = None
    timesteps: Optional[torch.Tensor] = None  # [B], one shared value during sampling

    def __len__(self) -> int:
        return len(self.X)

    @property
    def device(self) -> torch.device:
        return self.X.device
