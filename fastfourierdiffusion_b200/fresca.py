"""FreSca frequency scaling of the score -- mirror of ``fdiff.utils.fresca`` for the sampling path.

``apply_fresca_to_score`` (fresca.py:213-268) is what the sampler calls per step (sampler.py:79-93):
``rfft`` along the sequence axis, low band scaled by ``low_scale``, high band by a (mildly time-dependent)
``high_scale``, ``irfft``.  Here the two transforms are dense real-packed DFT matrix products and the band
cut-off -- including the data-dependent "energy" strategy, which the reference finds with a host loop of
``.item()`` calls (fresca.py:52-58) -- is decided on the device (``ffd_fresca``), so the filter adds no
host synchronisation to the step.  Only the 3-D ``(batch, seq_len, channels)`` case used by the sampler is built.
"""

from __future__ import annotations

from typing import Literal, Optional

import torch

from . import _lib
from .fourier import transform_matrix


def frequency_scale(x: torch.Tensor, low_scale: float = 1.0, high_scale: float = 1.0, cutoff_ratio: float = 0.5,
                    cutoff_strategy: Literal["spatial", "energy"] = "spatial", dim: int = 1) -> torch.Tensor:
    """fresca.py:111-210 for ``x[B, L, C]``."""
    if low_scale == 1.0 and high_scale == 1.0:
        return x
    if x.dim() != 3 or dim != 1:
        raise NotImplementedError("only the (batch, seq_len, channels) case of the sampling path is built")
    if cutoff_strategy not in ("spatial", "energy"):
        raise ValueError(f"Unknown cutoff_strategy: {cutoff_strategy}")
    dev = _lib.require_cuda()
    src = x.device
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, L, C = xd.shape
    lib = _lib.load()
    U, V = transform_matrix(L, False, dev), transform_matrix(L, True, dev)
    scratch = torch.empty(lib.ffd_fresca_scratch_floats(B, L, C), device=dev, dtype=torch.float32)
    out = torch.empty_like(xd)
    _lib.check(lib.ffd_fresca(_lib.ptr(U), _lib.ptr(V), _lib.ptr(xd), _lib.ptr(out), _lib.ptr(scratch), B, L, C,
                              float(low_scale), float(high_scale), float(cutoff_ratio),
                              1 if cutoff_strategy == "energy" else 0, _lib.stream_handle(dev)))
    return out if src.type == "cuda" else out.to(src)


def dynamic_high_scale(high_scale: float, timestep: Optional[float], num_steps: Optional[int]) -> float:
    """fresca.py:243-256: h(t) = (1 - t/num_steps) (h - 1) + 1 when h > 1 and both are given."""
    if timestep is not None and num_steps is not None and high_scale > 1.0:
        t_normalized = timestep / num_steps if num_steps > 0 else 0.0
        return (1.0 - t_normalized) * (high_scale - 1.0) + 1.0
    return high_scale


def apply_fresca_to_score(score: torch.Tensor, low_scale: float = 1.0, high_scale: float = 1.0,
                          cutoff_ratio: float = 0.5, cutoff_strategy: Literal["spatial", "energy"] = "energy",
                          timestep: Optional[float] = None, num_steps: Optional[int] = None) -> torch.Tensor:
    """fresca.py:213-268."""
    return frequency_scale(score, low_scale=low_scale, high_scale=dynamic_high_scale(high_scale, timestep, num_steps),
                           cutoff_ratio=cutoff_ratio, cutoff_strategy=cutoff_strategy)
