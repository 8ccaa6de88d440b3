"""Sample-quality metrics on the device -- host mirror of ``fdiff.sampling.metrics`` (SlicedWasserstein,
MarginalWasserstein; metrics.py:101-216) and ``fdiff.utils.wasserstein.WassersteinDistances`` (wasserstein.py:12-199).

Same constructor arguments, same result keys, same random directions (``np.random.default_rng(seed).normal`` per
direction, normalised: wasserstein.py:42-60).  The reference projects with numpy and calls POT's ``ot.emd2_1d`` per
direction on the host; here projections, sorts and the exact 1-D transport costs run in ``libffd_b200.so``
(``ffd_metric_*``) and only the ``n_directions`` distances come back.  Arithmetic: fp32 projections (the reference:
float64), fp64 accumulation of the transport cost.  No CPU fallback.
"""

from __future__ import annotations

from typing import Any, Optional, Union

import numpy as np
import torch

from . import _lib

ArrayLike = Union[np.ndarray, torch.Tensor]
MAX_ROWS = 32768  # samples per set (one sorted row lives in shared memory)


def _flat_device(x: ArrayLike) -> torch.Tensor:
    """check_flat_array (utils/tensors.py:5-22): (n, ...) -> (n, d), on the GPU as contiguous fp32."""
    dev = _lib.require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x.detach()
    if t.ndim > 2:
        t = t.reshape(t.shape[0], -1)
    assert t.ndim == 2, f"x must be a 2d array. Got {t.ndim}d array."
    return t.to(device=dev, dtype=torch.float32).contiguous()


class WassersteinDistances:
    """wasserstein.py:12-199 with the numpy / POT work moved to the device."""

    def __init__(self, original_data: ArrayLike, other_data: ArrayLike, normalisation: Optional[str] = "none",
                 seed: Optional[int] = None) -> None:
        self.original_data = _flat_device(original_data)
        self.other_data = _flat_device(other_data)
        if self.original_data.shape[1] != self.other_data.shape[1]:
            raise ValueError("original_data and other_data must have the same number of features")
        for t in (self.original_data, self.other_data):
            if t.shape[0] > MAX_ROWS:
                raise _lib.FFDError(f"at most {MAX_ROWS} samples per set are supported (got {t.shape[0]})")
        if normalisation not in ("none", "standardise"):
            raise ValueError(f"Unrecognised normalisation type: {normalisation}")
        self.normalisation = normalisation
        self.rng = np.random.default_rng(seed)

    # -- directions: host RNG, identical to the reference ---------------------------------------------
    def random_direction(self, dim: int) -> np.ndarray:
        vector = self.rng.normal(size=dim)
        return vector / np.linalg.norm(vector)

    def get_random_directions(self, n_directions: int) -> list:
        dimension = self.original_data.shape[1]
        return [self.random_direction(dimension) for _ in range(n_directions)]

    def get_marginal_directions(self) -> list:
        dimension = self.original_data.shape[1]
        return [np.identity(dimension)[i] for i in range(dimension)]

    # -- device work -------------------------------------------------------------------------------------
    def _distances(self, directions: np.ndarray) -> np.ndarray:
        """sqrt(emd2_1d(project(original), project(other))) for every row of ``directions`` [R, D]."""
        lib, st = _lib.load(), _lib.stream_handle()
        dev = self.original_data.device
        dirs = torch.from_numpy(np.ascontiguousarray(directions, dtype=np.float32)).to(dev)
        R, D = dirs.shape
        N, M = self.original_data.shape[0], self.other_data.shape[0]
        pa = torch.empty(R, N, dtype=torch.float32, device=dev)
        pb = torch.empty(R, M, dtype=torch.float32, device=dev)
        _lib.check(lib.ffd_metric_project(_lib.ptr(self.original_data), N, D, _lib.ptr(dirs), R, _lib.ptr(pa), st))
        _lib.check(lib.ffd_metric_project(_lib.ptr(self.other_data), M, D, _lib.ptr(dirs), R, _lib.ptr(pb), st))
        sd = None
        if self.normalisation == "standardise":  # wasserstein.py:157-159: both divided by np.std(original)
            sd = torch.empty(R, dtype=torch.float32, device=dev)
            _lib.check(lib.ffd_metric_row_std(_lib.ptr(pa), R, N, _lib.ptr(sd), st))
        _lib.check(lib.ffd_metric_sort_rows(_lib.ptr(pa), R, N, st))
        _lib.check(lib.ffd_metric_sort_rows(_lib.ptr(pb), R, M, st))
        out = torch.empty(R, dtype=torch.float32, device=dev)
        _lib.check(lib.ffd_metric_w2_sorted(_lib.ptr(pa), N, _lib.ptr(pb), M, R, _lib.ptr(out), st))
        if sd is not None:
            out = out / sd  # W2 is positively homogeneous: W2(a / s, b / s) = W2(a, b) / s
        return out.cpu().numpy().astype(np.float64)

    def feature_distance(self, feature: int) -> float:
        e = np.zeros((1, self.original_data.shape[1]))
        e[0, feature] = 1.0
        return float(self._distances(e)[0])

    def directional_distance(self, direction: np.ndarray) -> float:
        return float(self._distances(np.asarray(direction, dtype=np.float64)[None, :])[0])

    def sliced_distances(self, num_directions: int) -> np.ndarray:
        return self._distances(np.stack(self.get_random_directions(num_directions)))

    def marginal_distances(self) -> np.ndarray:
        return self._distances(np.identity(self.original_data.shape[1]))


class _Metric:
    def __init__(self, original_samples: ArrayLike) -> None:
        self.original_samples = _flat_device(original_samples)

    @property
    def baseline_metrics(self) -> dict:
        return {}


class SlicedWasserstein(_Metric):
    """metrics.py:101-158."""

    def __init__(self, original_samples: ArrayLike, random_seed: int, num_directions: int,
                 save_all_distances: bool = False) -> None:
        super().__init__(original_samples)
        self.random_seed = random_seed
        self.num_directions = num_directions
        self.save_all_distances = save_all_distances

    def __call__(self, other_samples: ArrayLike) -> dict:
        wd = WassersteinDistances(self.original_samples, other_samples, seed=self.random_seed)
        distances = wd.sliced_distances(self.num_directions)
        metrics: dict = {"sliced_wasserstein_mean": float(np.mean(distances)),
                         "sliced_wasserstein_max": float(np.max(distances))}
        if self.save_all_distances:
            metrics["sliced_wasserstein_all"] = distances.tolist()
        return metrics

    @property
    def baseline_metrics(self) -> dict:
        n = self.original_samples.shape[0]
        d_self = WassersteinDistances(self.original_samples[: n // 2], self.original_samples[n // 2:],
                                      seed=self.random_seed).sliced_distances(self.num_directions)
        avg = self.original_samples.mean(dim=0, keepdim=True)
        d_dummy = WassersteinDistances(self.original_samples, avg, seed=self.random_seed).sliced_distances(
            self.num_directions)
        return {"sliced_wasserstein_mean_self": float(np.mean(d_self)),
                "sliced_wasserstein_max_self": float(np.max(d_self)),
                "sliced_wasserstein_mean_dummy": float(np.mean(d_dummy)),
                "sliced_wasserstein_max_dummy": float(np.max(d_dummy))}

    @property
    def name(self) -> str:
        return "sliced_wasserstein"


class MarginalWasserstein(_Metric):
    """metrics.py:161-216."""

    def __init__(self, original_samples: ArrayLike, random_seed: int, save_all_distances: bool = False) -> None:
        super().__init__(original_samples)
        self.random_seed = random_seed
        self.save_all_distances = save_all_distances

    def __call__(self, other_samples: ArrayLike) -> dict:
        wd = WassersteinDistances(self.original_samples, other_samples, seed=self.random_seed)
        distances = wd.marginal_distances()
        metrics: dict = {"marginal_wasserstein_mean": float(np.mean(distances)),
                         "marginal_wasserstein_max": float(np.max(distances))}
        if self.save_all_distances:
            metrics["marginal_wasserstein_all"] = distances.tolist()
        return metrics

    @property
    def baseline_metrics(self) -> dict:
        n = self.original_samples.shape[0]
        d_self = WassersteinDistances(self.original_samples[: n // 2], self.original_samples[n // 2:],
                                      seed=self.random_seed).marginal_distances()
        avg = self.original_samples.mean(dim=0, keepdim=True)
        d_dummy = WassersteinDistances(self.original_samples, avg, seed=self.random_seed).marginal_distances()
        return {"marginal_wasserstein_mean_self": float(np.mean(d_self)),
                "marginal_wasserstein_max_self": float(np.max(d_self)),
                "marginal_wasserstein_mean_dummy": float(np.mean(d_dummy)),
                "marginal_wasserstein_max_dummy": float(np.max(d_dummy))}

    @property
    def name(self) -> str:
        return "marginal_wasserstein"


class MetricCollection:
    """metrics.py:28-98: every (partially instantiated) metric is evaluated on the time-domain samples and on their
    DFT; optionally the marginal Wasserstein distance between spectral densities; baselines merged in."""

    def __init__(self, metrics: list, original_samples: Optional[ArrayLike] = None, include_baselines: bool = True,
                 include_spectral_density: bool = False) -> None:
        from functools import partial

        from .fourier import dft, spectral_density

        self._dft, self._spectral = dft, spectral_density
        metrics_time, metrics_freq = [], []
        orig = None if original_samples is None else self._as_tensor(original_samples)
        orig_freq = dft(orig) if orig is not None else None
        for metric in metrics:
            if isinstance(metric, partial):  # metrics.py:45-51
                assert orig is not None, "Original samples must be provided for the metrics to be instantiated."
                metrics_time.append(metric(original_samples=orig))
                metrics_freq.append(metric(original_samples=orig_freq))
        self.metrics_time, self.metrics_freq = metrics_time, metrics_freq
        self.include_baselines = include_baselines
        self.metric_spectral = (MarginalWasserstein(original_samples=spectral_density(orig), random_seed=42,
                                                    save_all_distances=True)
                                if include_spectral_density else None)

    @staticmethod
    def _as_tensor(x: ArrayLike) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x.detach()
        return t.to(device=_lib.require_cuda(), dtype=torch.float32)

    def __call__(self, other_samples: ArrayLike) -> dict:
        other = self._as_tensor(other_samples)
        other_freq = self._dft(other)
        out: dict = {}
        for mt, mf in zip(self.metrics_time, self.metrics_freq):
            out.update({f"time_{k}": v for k, v in mt(other).items()})
            out.update({f"freq_{k}": v for k, v in mf(other_freq).items()})
        if self.include_baselines:
            out.update(self.baseline_metrics)
        if self.metric_spectral is not None:
            out.update({f"spectral_{k}": v for k, v in self.metric_spectral(self._spectral(other)).items()})
        return dict(sorted(out.items(), key=lambda item: item[0]))

    @property
    def baseline_metrics(self) -> dict:
        out: dict = {}
        for mt, mf in zip(self.metrics_time, self.metrics_freq):
            out.update({f"time_{k}": v for k, v in mt.baseline_metrics.items()})
            out.update({f"freq_{k}": v for k, v in mf.baseline_metrics.items()})
        return out
