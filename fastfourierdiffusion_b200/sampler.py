"""Reverse-SDE sampler -- host mirror of ``fdiff.sampling.sampler.DiffusionSampler``.

Same constructor and ``sample`` / ``reverse_diffusion_step`` / ``sample_prior``
contract as the reference (sampler.py:14-228).  The step loop itself
(sampler.py:157-210) is one ``ffd_sample_steps`` call per noise chunk: the
recompute schedule, the per-step SDE scalars and the time-embedding table are
precomputed (they depend on the step index only), so the device runs all steps
back to back with no host synchronisation, where the reference synchronises
three times per step.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Literal, Optional, Set

import torch

from . import _lib
from .batch import DiffusableBatch
from .score_model import ScoreModule
from .sde import SDE


class DiffusionSampler:
    def __init__(self, score_model: ScoreModule, sample_batch_size: int, use_cache: bool = False,
                 cache_kwargs: Optional[dict] = None, use_fresca: bool = False, fresca_low_scale: float = 1.0,
                 fresca_high_scale: float = 1.5, fresca_cutoff_ratio: float = 0.5,
                 fresca_cutoff_strategy: Literal["spatial", "energy"] = "energy",
                 noise: Literal["torch", "philox"] = "torch", seed: int = 0) -> None:
        self.score_model = score_model
        self.noise_scheduler = score_model.noise_scheduler
        self.sample_batch_size = sample_batch_size
        self.n_channels = score_model.n_channels
        self.max_len = score_model.max_len
        self.use_cache = use_cache
        if use_cache:
            self.score_model.enable_caching(**(cache_kwargs or {}))
        self.use_fresca = use_fresca
        self.fresca_low_scale = fresca_low_scale
        self.fresca_high_scale = fresca_high_scale
        self.fresca_cutoff_ratio = fresca_cutoff_ratio
        self.fresca_cutoff_strategy = fresca_cutoff_strategy
        # "torch": z from torch.randn / randn_like on the same generators the reference uses
        # (CPU for the prior, device for the steps: sde.py:85,157,238); "philox": drawn inside
        # the step kernel, keyed (seed, batch, step).
        assert noise in ("torch", "philox")
        self.noise = noise
        self.seed = seed
        self.noise_chunk_steps = 8  # steps of torch noise generated per ffd_sample_steps call

    # -- single step (API parity; the fused loop below does not go through here) --------
    def reverse_diffusion_step(self, batch: DiffusableBatch, step: int = 0,
                               recompute_tokens: Optional[Set[int]] = None,
                               noise: Optional[torch.Tensor] = None, t: Optional[float] = None) -> torch.Tensor:
        """sampler.py:48-103.  ``noise`` (extra, optional) injects z instead of drawing ``randn_like``; ``t`` (extra,
        optional) is the batch's shared timestep as a host float."""
        X, timesteps = batch.X, batch.timesteps
        assert timesteps is not None and timesteps.size(0) == len(batch)
        model = self.score_model
        # One shared t per batch (sampler.py:59-60).  The loops of this class pass it as ``t`` so that the three
        # ``.item()`` reads the reference makes per step (device syncs when ``timesteps`` is a CUDA tensor) do not
        # happen; a caller that only has the tensor pays for one read.
        if t is None:
            assert torch.min(timesteps) == torch.max(timesteps)
            t = float(timesteps[0].item())
        if self.use_cache and recompute_tokens is not None:
            score, crf = model(batch, recompute_tokens=recompute_tokens, step=step, return_crf=True, t=t)
            if model.cache is not None:
                model.cache.update_crf(crf, timestep=t)
                model.cache.current_step = step
        else:
            score = model(batch, t=t)
        if self.use_fresca:  # sampler.py:79-93
            from .fresca import apply_fresca_to_score

            strategy = "energy" if self.fresca_cutoff_strategy == "energy" else "spatial"
            score = apply_fresca_to_score(score, low_scale=self.fresca_low_scale, high_scale=self.fresca_high_scale,
                                          cutoff_ratio=self.fresca_cutoff_ratio, cutoff_strategy=strategy,
                                          timestep=t,
                                          num_steps=getattr(self, "_num_diffusion_steps", None))
        out = self.noise_scheduler.step(model_output=score, timestep=t, sample=X, noise=noise)
        return out.prev_sample

    # -- the loop ---------------------------------------------------------------------------
    def sample(self, num_samples: int, num_diffusion_steps: Optional[int] = None) -> torch.Tensor:
        """sampler.py:105-215: returns a CPU fp32 tensor [floor(N/bs)*bs, L, C] (one batch of N if N < bs)."""
        return self._sample(num_samples, num_diffusion_steps, None)

    def sample_time_domain(self, num_samples: int, num_diffusion_steps: Optional[int] = None,
                           feature_std: Optional[torch.Tensor] = None,
                           feature_mean: Optional[torch.Tensor] = None,
                           fourier_transform: Optional[bool] = None) -> torch.Tensor:
        """``sample`` followed by the tail of cmd/sample.py:107-113 (de-standardise ``X * std + mean``, then ``idft``
        for a frequency-domain model) ON THE DEVICE, per batch, before the copy to the host: only time-domain
        samples cross PCIe.  Same values as ``idft(sample(...), std, mean)``.
        ``fourier_transform`` (cfg.fourier_transform of cmd/sample.py) defaults to the model's ``scale_noise``."""
        from .fourier import idft

        freq = bool(self.score_model.scale_noise) if fourier_transform is None else bool(fourier_transform)

        def tail(X: torch.Tensor) -> torch.Tensor:
            if freq:
                return idft(X, feature_std, feature_mean)
            if feature_std is not None:
                X = X * feature_std.to(X.device)
            if feature_mean is not None:
                X = X + feature_mean.to(X.device)
            return X

        return self._sample(num_samples, num_diffusion_steps, tail)

    def sample_on_device(self, num_samples: int, num_diffusion_steps: Optional[int] = None) -> torch.Tensor:
        """``sample`` without the final copy to the host: the ``[N, L, C]`` result stays on the model's device (what
        a rank of a batch-sharded run hands to the final NCCL gather, ``distributed.sample_sharded``)."""
        return self._sample(num_samples, num_diffusion_steps, None, to_host=False)

    def _sample(self, num_samples: int, num_diffusion_steps: Optional[int], tail, to_host: bool = True) -> torch.Tensor:
        if num_diffusion_steps is not None:
            self._num_diffusion_steps = num_diffusion_steps
        model = self.score_model
        model.eval()
        n_steps = model.num_training_steps if num_diffusion_steps is None else num_diffusion_steps
        self.noise_scheduler.set_timesteps(n_steps)
        all_samples: List[torch.Tensor] = []
        num_batches = max(1, num_samples // self.sample_batch_size)
        global_step = 0
        with torch.no_grad():
            for batch_idx in range(num_batches):
                batch_size = min(num_samples - batch_idx * self.sample_batch_size, self.sample_batch_size)
                X = self.sample_prior(batch_size, batch_idx)
                if self.use_cache and model.cache is not None and batch_idx == 0:
                    model.cache.reset()
                    global_step = 0
                X = self.run_steps(X, global_step, batch_idx=batch_idx)
                global_step += n_steps
                if tail is not None:
                    X = tail(X)
                all_samples.append(X.cpu() if to_host else X)
        return torch.cat(all_samples, dim=0)

    def sample_prior(self, batch_size: int, batch_idx: int = 0) -> torch.Tensor:
        """sampler.py:217-228."""
        if not isinstance(self.noise_scheduler, SDE):
            raise NotImplementedError("Scheduler not recognized.")
        shape = (batch_size, self.max_len, self.n_channels)
        if self.noise == "torch":
            return self.noise_scheduler.prior_sampling(shape).to(device=self.score_model.device)
        return self.noise_scheduler.prior_from_noise(None, shape, seed=self.seed,
                                                     stream_id=self._stream_base(batch_idx))

    def _stream_base(self, batch_idx: int) -> int:
        # one Philox stream per (batch, step): stream 0 of a batch is its prior
        return batch_idx * (1 << 32)

    def run_steps_unfused(self, X: torch.Tensor, global_step0: int = 0,
                          noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The reference's own per-step loop (sampler.py:157-210) over ``reverse_diffusion_step``: one Python
        iteration per step over the same kernels.  ``sample`` does not use it (FreSca included, the loop is the fused
        ``ffd_sample_steps`` call of ``run_steps``); it is kept as the step-by-step form of the public API and as the
        cross-check of the fused loop in the tests.  The shared timestep is passed as a host float, so there is no
        ``.item()`` on a device tensor."""
        model = self.score_model
        dev = model.device
        X = X.to(device=dev, dtype=torch.float32).contiguous()
        cache = model.cache if (self.use_cache and model.cache is not None) else None
        if cache is not None and getattr(cache, "event_driven", False):
            raise NotImplementedError("the device-decided event schedule (event_driven=True) runs inside the fused "
                                      "step loop (ffd_sample_steps); it is not available together with use_fresca")
        for step_idx, t in enumerate(self.noise_scheduler.timesteps.tolist()):
            timesteps = torch.full((X.shape[0],), t, dtype=torch.float32, device=dev)
            batch = DiffusableBatch(X=X, y=None, timesteps=timesteps)
            recompute = None
            gs = global_step0 + step_idx
            if cache is not None:
                cache.current_step = gs
                recompute = cache.determine_recompute_set(x_tilde=None, event_intensity=0.1, step=gs)
            X = self.reverse_diffusion_step(batch, step=step_idx, recompute_tokens=recompute,
                                            noise=None if noise is None else noise[step_idx], t=t)
        return X

    def step_tables(self, global_step0: int):
        """Per-step host tables for one batch: SDE scalars, |recompute set|, CRF-store flag."""
        sched = self.noise_scheduler
        model = self.score_model
        ts = [float(t) for t in sched.timesteps.tolist()]
        n = len(ts)
        coef = (_lib.StepCoef * n)(*[sched.step_coef(t) for t in ts])
        n_own = (C.c_int32 * n)()
        crf = (C.c_int32 * n)()
        cache = model.cache if (self.use_cache and model.cache is not None) else None
        for i in range(n):
            if cache is None:
                n_own[i] = -1
            else:
                gs = global_step0 + i
                # x_tilde / event_intensity are ignored by the reference's schedule
                n_own[i] = len(cache.determine_recompute_set(None, 0.1, gs))
                crf[i] = 1 if cache.crf_stored_at(gs) else 0
        return coef, n_own, crf

    def run_steps(self, X: torch.Tensor, global_step0: int = 0, batch_idx: int = 0,
                  noise: Optional[torch.Tensor] = None, step_begin: int = 0,
                  step_end: Optional[int] = None, noise_step0: int = 0) -> torch.Tensor:
        """Diffusion steps [step_begin, step_end) of one batch on the device.  ``noise``
        ([k, B, L, C], holding steps noise_step0 .. noise_step0+k) injects the per-step z
        (parity runs); otherwise z follows ``self.noise``."""
        model = self.score_model
        sched = self.noise_scheduler
        dev = model.device
        lib = _lib.load()
        B = X.shape[0]
        n = sched.timesteps.numel()
        # Launch-bound regime (few token tiles): runs of pure-cache steps are replayed as a CUDA graph by the library.
        # A graph is keyed on its buffers, so the state / noise / time-embedding buffers persist across calls then.
        env = os.environ.get("FFD_GRAPH")
        use_graph = (env == "1") if env in ("0", "1") else (B * self.max_len <= 131072)
        if use_graph:
            key = (B, n, str(dev))
            if getattr(self, "_gkey", None) != key:
                shape = (B, self.max_len, self.n_channels)
                self._gx = torch.empty(shape, device=dev, dtype=torch.float32)
                self._gx_tmp = torch.empty_like(self._gx)
                self._gnoise = None
                self._gscratch = torch.empty(lib.ffd_graph_scratch_bytes(n), device=dev, dtype=torch.uint8)
                self._gte = None
                self._gkey = key
            self._gx.copy_(X.to(device=dev, dtype=torch.float32))
            X = self._gx
        else:
            X = X.to(device=dev, dtype=torch.float32).contiguous().clone()
        step_end = n if step_end is None else step_end
        coef, n_own, crf_flags = self.step_tables(global_step0)
        te_key = (n, model._arena_key, float(sched.timesteps[0]), float(sched.timesteps[-1]))
        if use_graph and self._gte is not None and self._gte[0] == te_key:
            te_table = self._gte[1]
        else:
            te_table = model.time_embedding(sched.timesteps)
            if use_graph:
                self._gte = (te_key, te_table)
        a = _lib.SampleArgs()
        a.fwd = model.forward_args(B)
        cache = model.cache if (self.use_cache and model.cache is not None) else None
        if cache is not None:
            a.fwd.crf_out = _lib.ptr(cache._crf_store)
        a.G = _lib.ptr(sched.G_device(dev))
        a.time_emb_table = _lib.ptr(te_table)
        a.coef_host = coef
        a.n_own_host = n_own
        a.store_crf_host = crf_flags
        a.seed = self.seed
        a.stream_base = self._stream_base(batch_idx) + 1
        event = None
        if cache is not None and cache.event_driven:
            # the device-decided schedule: K/V rows live with the first cache of the model
            event = cache.event_args(model.dims.d_model)
            a.event = C.pointer(event)
        fresca_keep = None
        if self.use_fresca:
            # FreSca between score and SDE update (sampler.py:79-93) inside the fused loop: the time-dependent
            # high-band scale is a per-step host table, both cut-off strategies are decided on the device
            from .fourier import transform_matrix
            from .fresca import dynamic_high_scale

            L, Cn = self.max_len, self.n_channels
            fr = _lib.FrescaArgs()
            U, V = transform_matrix(L, False, dev), transform_matrix(L, True, dev)
            scratch = torch.empty(lib.ffd_fresca_scratch_floats(B, L, Cn) + 2 * B * L * Cn, device=dev,
                                  dtype=torch.float32)
            num_steps = getattr(self, "_num_diffusion_steps", None)
            hs = (C.c_float * n)(*[dynamic_high_scale(self.fresca_high_scale, float(t), num_steps)
                                   for t in sched.timesteps.tolist()])
            fr.U, fr.V, fr.scratch = _lib.ptr(U), _lib.ptr(V), _lib.ptr(scratch)
            fr.high_scale_host = hs
            fr.low_scale = float(self.fresca_low_scale)
            fr.cutoff_ratio = float(self.fresca_cutoff_ratio)
            fr.strategy = 1 if self.fresca_cutoff_strategy == "energy" else 0
            a.fresca = C.pointer(fr)
            fresca_keep = (fr, U, V, scratch, hs)
        x_tmp = self._gx_tmp if use_graph else torch.empty_like(X)
        if use_graph:
            a.graph = 1
            a.graph_scratch = _lib.ptr(self._gscratch)
            a.graph_scratch_bytes = self._gscratch.numel()
        a.x = _lib.ptr(X)
        a.x_tmp = _lib.ptr(x_tmp)
        st = _lib.stream_handle(dev)
        if noise is not None:
            noise = noise.to(device=dev, dtype=torch.float32).contiguous()
            assert tuple(noise.shape[1:]) == tuple(X.shape)
            assert noise_step0 <= step_begin and step_end - noise_step0 <= noise.shape[0]
            a.noise = _lib.ptr(noise)
            a.noise_step0 = noise_step0
            a.step_begin, a.step_end = step_begin, step_end
            _lib.check(lib.ffd_sample_steps(a, st))
        elif self.noise == "philox":
            a.noise = None
            a.step_begin, a.step_end = step_begin, step_end
            _lib.check(lib.ffd_sample_steps(a, st))
        else:
            # torch noise: same generator stream as the reference's randn_like per step
            chunk = max(1, self.noise_chunk_steps)
            if use_graph:
                if self._gnoise is None or self._gnoise.shape[0] != chunk:
                    self._gnoise = torch.empty((chunk,) + tuple(X.shape), device=dev, dtype=torch.float32)
                buf = self._gnoise
            else:
                buf = torch.empty((chunk,) + tuple(X.shape), device=dev, dtype=torch.float32)
            for s0 in range(step_begin, step_end, chunk):
                s1 = min(step_end, s0 + chunk)
                for j in range(s1 - s0):
                    buf[j].normal_()
                a.noise = _lib.ptr(buf)
                a.noise_step0 = s0
                a.step_begin, a.step_end = s0, s1
                _lib.check(lib.ffd_sample_steps(a, st))
        del fresca_keep
        if cache is not None:
            owner = model._kv_owner
            for i in range(step_begin, step_end):
                owner.account_forward(int(n_own[i]))
                if crf_flags[i]:
                    cache._crf_valid = True
            cache.current_step = step_end - 1 if step_end > step_begin else cache.current_step
        return X.clone() if use_graph else X
