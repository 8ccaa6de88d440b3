"""VP / VE SDE schedulers -- host mirror of ``fdiff.schedulers.sde``.

Only the sampling half of the reference class is on the hot path
(src/fdiff/schedulers/sde.py:42-87 set-up and prior, :129-165 VE step,
:212-246 VP step); ``marginal_prob`` / ``add_noise`` are training code and are
deliberately absent.  The per-step scalars are evaluated on the host in Python
floats exactly as the reference does, the element-wise update runs in
``ffd_sde_step`` / the fused step epilogue on the device.
"""

from __future__ import annotations

import math
from collections import namedtuple
from typing import Optional, Tuple

import torch

from . import _lib

SamplingOutput = namedtuple("SamplingOutput", ["prev_sample"])


class SDE:
    """Common part: noise scaling G, time grid, prior (sde.py:14-87)."""

    kind = -1

    def __init__(self, fourier_noise_scaling: bool = False, eps: float = 1e-5):
        self.noise_scaling = fourier_noise_scaling
        self.eps = eps
        self.G: Optional[torch.Tensor] = None
        self._G_dev: Optional[torch.Tensor] = None

    @property
    def T(self) -> float:
        return 1.0

    # -- set-up -----------------------------------------------------------
    def set_noise_scaling(self, max_len: int) -> None:
        """G = 1 (time domain) or 1/sqrt(2) with the DC (and Nyquist) entries 1 (sde.py:42-60)."""
        G = torch.ones(max_len)
        if self.noise_scaling:
            G = 1 / (math.sqrt(2)) * G
            G[0] *= math.sqrt(2)
            if max_len % 2 == 0:
                G[max_len // 2] *= math.sqrt(2)
        self.G = G
        self.G_matrix = torch.diag(G)  # kept for API compatibility; never multiplied
        self._G_dev = None

    def set_timesteps(self, num_diffusion_steps: int) -> None:
        """sde.py:62-64.  A single step has no step size, as in the reference (IndexError)."""
        self.timesteps = torch.linspace(1.0, self.eps, num_diffusion_steps)
        self.step_size = self.timesteps[0] - self.timesteps[1]

    def G_device(self, device: torch.device) -> torch.Tensor:
        assert self.G is not None, "call set_noise_scaling(max_len) first"
        cached = getattr(self, "_G_dev", None)  # absent on a scheduler unpickled from a reference checkpoint
        if cached is None or cached.device != device:
            self._G_dev = self.G.to(device=device, dtype=torch.float32).contiguous()
        return self._G_dev

    # -- prior ---------------------------------------------------------------
    prior_scale = 1.0

    def prior_sampling(self, shape: Tuple[int, ...]) -> torch.Tensor:
        """G z with z ~ N(0, I) drawn from the CPU generator like the reference
        (sde.py:79-87); the product G z (a diagonal matmul there) runs on the device."""
        dev = _lib.require_cuda()
        assert len(shape) == 3 and self.G is not None and shape[1] == self.G.shape[0]
        z = torch.randn(*shape).to(dev)
        return self.prior_from_noise(z)

    def prior_from_noise(self, z: Optional[torch.Tensor], shape=None, seed: int = 0, stream_id: int = 0):
        """z given -> G z; z None -> in-kernel Philox draw keyed (seed, stream_id)."""
        dev = _lib.require_cuda()
        if z is not None:
            shape = tuple(z.shape)
            z = z.to(device=dev, dtype=torch.float32).contiguous()
        B, L, C = shape
        out = torch.empty(shape, device=dev, dtype=torch.float32)
        _lib.check(_lib.load().ffd_prior(_lib.ptr(self.G_device(dev)), _lib.ptr(z), seed, stream_id,
                                         float(self.prior_scale), _lib.ptr(out), B, L, C,
                                         _lib.stream_handle(dev)))
        return out

    # -- reverse step -----------------------------------------------------------
    def step_coef(self, timestep: float) -> "_lib.StepCoef":
        raise NotImplementedError

    def step(self, model_output: torch.Tensor, timestep: float, sample: torch.Tensor,
             noise: Optional[torch.Tensor] = None) -> SamplingOutput:
        """One Euler-Maruyama reverse step.  ``noise`` None draws ``torch.randn_like``
        like the reference (sde.py:157,238)."""
        assert self.step_size > 0
        dev = _lib.require_cuda()
        x = sample.to(device=dev, dtype=torch.float32).contiguous()
        s = model_output.to(device=dev, dtype=torch.float32).contiguous()
        z = torch.randn_like(x) if noise is None else noise.to(device=dev, dtype=torch.float32).contiguous()
        B, L, C = x.shape
        out = torch.empty_like(x)
        coef = self.step_coef(timestep)
        _lib.check(_lib.load().ffd_sde_step(coef, _lib.ptr(self.G_device(dev)), _lib.ptr(x), _lib.ptr(s),
                                            _lib.ptr(z), 0, 0, _lib.ptr(out), B, L, C, _lib.stream_handle(dev)))
        return SamplingOutput(prev_sample=out)

    def _dt(self) -> Tuple[float, float]:
        return float(self.step_size.item()), float(torch.sqrt(self.step_size).item())


class VEScheduler(SDE):
    kind = _lib.SDE_VE

    def __init__(self, sigma_min: float = 0.01, sigma_max: float = 50.0, fourier_noise_scaling: bool = False,
                 eps: float = 1e-5):
        super().__init__(fourier_noise_scaling=fourier_noise_scaling, eps=eps)
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max

    @property
    def prior_scale(self) -> float:  # sde.py:125-127
        return self.sigma_max

    def step_coef(self, timestep: float) -> "_lib.StepCoef":
        # sde.py:143-147
        sd = (self.sigma_min * math.sqrt(2 * math.log(self.sigma_max / self.sigma_min))
              * (self.sigma_max / self.sigma_min) ** timestep)
        dt, sdt = self._dt()
        return _lib.StepCoef(self.kind, 0.0, sd, dt, sdt)


class VPScheduler(SDE):
    kind = _lib.SDE_VP

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0, fourier_noise_scaling: bool = False,
                 eps: float = 1e-5):
        super().__init__(fourier_noise_scaling=fourier_noise_scaling, eps=eps)
        self.beta_0 = beta_min
        self.beta_1 = beta_max

    def get_beta(self, timestep: float) -> float:  # sde.py:212-213
        return self.beta_0 + timestep * (self.beta_1 - self.beta_0)

    def step_coef(self, timestep: float) -> "_lib.StepCoef":
        beta = self.get_beta(timestep)
        dt, sdt = self._dt()
        return _lib.StepCoef(self.kind, -0.5 * beta, math.sqrt(beta), dt, sdt)
