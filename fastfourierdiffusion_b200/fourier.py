"""Boundary transforms ``dft`` / ``idft`` -- mirror of ``fdiff.utils.fourier`` (fourier.py:8-94).

The real-packed orthonormal DFT of a short, non-power-of-two series is a dense
L x L real matrix; ``ffd_dft_matrix`` builds it on the device
(entries evaluated in fp64, stored fp32) and ``ffd_dft_apply`` contracts it with
the batch.  Input may live on the CPU (as in the reference flows) -- it is
moved to the GPU, transformed there, and returned on its original device.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib

_MATRICES: Dict[Tuple[int, int, int], torch.Tensor] = {}


def transform_matrix(max_len: int, inverse: bool, device: torch.device) -> torch.Tensor:
    key = (max_len, int(inverse), device.index or 0)
    if key not in _MATRICES:
        T = torch.empty(max_len, max_len, device=device, dtype=torch.float32)
        _lib.check(_lib.load().ffd_dft_matrix(_lib.ptr(T), max_len, int(inverse), _lib.stream_handle(device)))
        _MATRICES[key] = T
    return _MATRICES[key]


def _apply(x: torch.Tensor, inverse: bool, scale: Optional[torch.Tensor] = None,
           shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _lib.require_cuda()
    if torch.is_complex(x):
        x = torch.real(x)
    assert x.dim() == 3, "expected (batch_size, max_len, n_channels)"
    src = x.device
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, L, Cc = xd.shape
    T = transform_matrix(L, inverse, dev)
    sc = None if scale is None else scale.to(device=dev, dtype=torch.float32).expand(L, Cc).contiguous()
    sh = None if shift is None else shift.to(device=dev, dtype=torch.float32).expand(L, Cc).contiguous()
    y = torch.empty_like(xd)
    _lib.check(_lib.load().ffd_dft_apply(_lib.ptr(T), _lib.ptr(xd), _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(y), B, L,
                                         Cc, _lib.stream_handle(dev)))
    return y if src.type == "cuda" else y.to(src)


def dft(x: torch.Tensor) -> torch.Tensor:
    """(B, L, C) time series -> real-packed orthonormal DFT of the same shape."""
    return _apply(x, inverse=False)


def idft(x: torch.Tensor, feature_std: Optional[torch.Tensor] = None,
         feature_mean: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Inverse of :func:`dft`.  ``feature_std`` / ``feature_mean`` ([L, C]) fuse the
    de-standardisation of cmd/sample.py:107-109 in front of the transform."""
    return _apply(x, inverse=True, scale=feature_std, shift=feature_mean)


def spectral_density(x: torch.Tensor, apply_dft: bool = True) -> torch.Tensor:
    """fourier.py:97-131: (B, L, C) time series (or, with ``apply_dft=False``, its real-packed DFT) ->
    |X_k|^2 of shape (B, L // 2 + 1, C)."""
    dev = _lib.require_cuda()
    assert x.dim() == 3, "expected (batch_size, max_len, n_channels)"
    src = x.device
    xt = (dft(x) if apply_dft else x).detach().to(device=dev, dtype=torch.float32).contiguous()
    B, L, Cc = xt.shape
    out = torch.empty(B, L // 2 + 1, Cc, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().ffd_spectral_density(_lib.ptr(xt), _lib.ptr(out), B, L, Cc, _lib.stream_handle(dev)))
    return out if src.type == "cuda" else out.to(src)
