"""Transformer score model -- host mirror of ``fdiff.models.score_models.ScoreModule``.

The module owns the same parameters under the same ``state_dict`` keys as the
reference (score_models.py:55-66: ``pos_encoder.embedding``, ``time_encoder.{W,dense}``,
``embedder``, ``unembedder``, ``backbone.layers.*``), created in the same order so
``torch.manual_seed(s)`` yields identical random-init weights, and a reference
checkpoint's ``state_dict`` loads unchanged.  The torch modules are parameter
containers only: ``forward`` flattens them into one fp32 device arena and runs the
CUDA path (``ffd_score_forward``); there is no eager/ATen fallback.

Training hooks of the LightningModule (losses, optimiser) are out of scope.
"""

from __future__ import annotations

import math
import os
from typing import Optional, Set, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from .batch import DiffusableBatch
from .caching import E2CRFCache
from .sde import SDE


class PositionalEncoding(nn.Module):
    """Learned table with max-norm sqrt(d) (transformer.py:8-29); the lookup + add runs in the embed kernel."""

    def __init__(self, d_model: int, max_len: int):
        super().__init__()
        self.embedding = nn.Embedding(num_embeddings=max_len, embedding_dim=d_model, max_norm=math.sqrt(d_model))


class GaussianFourierProjection(nn.Module):
    """Frozen Gaussian frequencies + dense layer (transformer.py:61-91); evaluated by ffd_time_embedding."""

    def __init__(self, d_model: int, scale: float = 30.0):
        super().__init__()
        self.d_model = d_model
        self.W = nn.Parameter(torch.randn((d_model + 1) // 2) * scale, requires_grad=False)
        self.dense = nn.Linear(d_model, d_model)


class _CachedBackboneMarker:
    """Stands where the reference keeps its ModuleList of cached layers (score_models.py:232-283);
    the cached branches are kernel modes here, so there is nothing to clone."""

    def __init__(self, num_layers: int):
        self.num_layers = num_layers

    def __len__(self) -> int:
        return self.num_layers


class ScoreModule(nn.Module):
    def __init__(self, n_channels: int, max_len: int, noise_scheduler: SDE, fourier_noise_scaling: bool = True,
                 d_model: int = 60, num_layers: int = 3, n_head: int = 12, num_training_steps: int = 1000,
                 lr_max: float = 1e-3, likelihood_weighting: bool = False, precision: Optional[str] = None) -> None:
        super().__init__()
        if not isinstance(noise_scheduler, SDE):
            raise NotImplementedError(f"Scheduler {noise_scheduler} not implemented yet, cannot set time encoder.")
        self.max_len = max_len
        self.n_channels = n_channels
        self.noise_scheduler = noise_scheduler
        self.num_warmup_steps = num_training_steps // 10
        self.num_training_steps = num_training_steps
        self.lr_max = lr_max
        self.d_model = d_model
        self.n_head = n_head
        self.scale_noise = fourier_noise_scaling
        self.likelihood_weighting = likelihood_weighting
        # same construction order as score_models.py:55-66 (RNG parity)
        self.pos_encoder = PositionalEncoding(d_model=d_model, max_len=max_len)
        self.time_encoder = GaussianFourierProjection(d_model=d_model)
        self.embedder = nn.Linear(in_features=n_channels, out_features=d_model)
        self.unembedder = nn.Linear(in_features=d_model, out_features=n_channels)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=n_head, batch_first=True)
        self.backbone = nn.TransformerEncoder(encoder_layer=layer, num_layers=num_layers)
        self.cached_backbone: Optional[_CachedBackboneMarker] = None
        self.cache: Optional[E2CRFCache] = None
        self.use_cache: bool = False
        self._kv_owner: Optional[E2CRFCache] = None  # first cache ever created
        self._precision = precision or os.environ.get("FFD_PRECISION")
        if self._precision == "bf16":  # former name of the 16-bit tensor-core path
            self._precision = "fp16"
        self._arena: Optional[torch.Tensor] = None
        self._arena_key = None
        self._packed: Optional[torch.Tensor] = None
        self._workspace: Optional[torch.Tensor] = None

    # -- plumbing ---------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def dims(self) -> "_lib.Dims":
        l0 = self.backbone.layers[0]
        return _lib.Dims(self.max_len, self.n_channels, self.d_model, l0.self_attn.num_heads,
                         l0.linear1.out_features, len(self.backbone.layers))

    @property
    def precision(self) -> str:
        """'fp16' (tcgen05 GEMMs on fp16 operands, fp32 accumulate and residual stream) when the kernels
        support the model shape, else 'fp32' (CUDA-core FMA).  'bf16' is accepted as the former name of 'fp16'."""
        if self._precision is None:
            lib = _lib.load()
            d = self.dims
            ok = lib.ffd_has_tensor_path() and lib.ffd_packed_bytes(d) > 0 and self._tc_ok()
            self._precision = "fp16" if ok else "fp32"
        return self._precision

    def set_precision(self, precision: str) -> None:
        assert precision in ("fp32", "fp16", "bf16")
        self._precision = "fp16" if precision == "bf16" else precision
        self._arena_key = None

    def _tc_ok(self) -> bool:
        d = self.dims
        return d.d_model + 2 <= 80 and d.d_ff % 128 == 0

    def _ordered_params(self):
        ps = [self.pos_encoder.embedding.weight, self.embedder.weight, self.embedder.bias,
              self.unembedder.weight, self.unembedder.bias]
        for lyr in self.backbone.layers:
            ps += [lyr.self_attn.in_proj_weight, lyr.self_attn.in_proj_bias, lyr.self_attn.out_proj.weight,
                   lyr.self_attn.out_proj.bias, lyr.linear1.weight, lyr.linear1.bias, lyr.linear2.weight,
                   lyr.linear2.bias, lyr.norm1.weight, lyr.norm1.bias, lyr.norm2.weight, lyr.norm2.bias]
        return ps

    def device_weights(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(fp32 arena, 16-bit tensor-core operand images or None); rebuilt when a parameter changed."""
        dev = self.device
        if dev.type != "cuda":
            raise _lib.FFDError("ScoreModule parameters must be on a CUDA device (no CPU fallback): call .to('cuda')")
        ps = self._ordered_params()
        key = (self.precision,) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._arena_key:
            lib = _lib.load()
            st = _lib.stream_handle(dev)
            pos = self.pos_encoder.embedding.weight
            with torch.no_grad():
                # nn.Embedding(max_norm) renorms the looked-up rows in place on every forward
                # (transformer.py:13-15,27): all L rows are looked up, so the parameter itself changes.
                _lib.check(lib.ffd_posemb_renorm(_lib.ptr(pos.data), self.max_len, self.d_model,
                                                 math.sqrt(self.d_model), st))
                self._arena = torch.cat([p.detach().reshape(-1).float() for p in ps]).contiguous()
            d = self.dims
            assert self._arena.numel() == lib.ffd_weights_count(d)
            self._packed = None
            if self.precision == "fp16":
                nbytes = lib.ffd_packed_bytes(d)
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.check(lib.ffd_pack_weights(d, _lib.ptr(self._arena), _lib.ptr(self._packed), st))
            self._arena_key = key
        return self._arena, self._packed

    def workspace(self, batch_size: int) -> torch.Tensor:
        need = _lib.load().ffd_workspace_bytes(self.dims, batch_size, self._prec_code())
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != self.device:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def _prec_code(self) -> int:
        return _lib.PREC_FP16 if self.precision == "fp16" else _lib.PREC_FP32

    def time_embedding(self, timesteps: torch.Tensor) -> torch.Tensor:
        """[n] -> [n, d]: dense([sin, cos](2 pi t W))[:d]  (transformer.py:80-91)."""
        dev = self.device
        t = timesteps.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(t.numel(), self.d_model, device=dev, dtype=torch.float32)
        te = self.time_encoder
        _lib.check(_lib.load().ffd_time_embedding(
            _lib.ptr(t), t.numel(), _lib.ptr(te.W.detach().contiguous()), _lib.ptr(te.dense.weight.detach().contiguous()),
            _lib.ptr(te.dense.bias.detach().contiguous()), self.d_model, _lib.ptr(out), _lib.stream_handle(dev)))
        return out

    def forward_args(self, B: int) -> "_lib.ForwardArgs":
        arena, packed = self.device_weights()
        ws = self.workspace(B)
        a = _lib.ForwardArgs()
        a.dims = self.dims
        a.precision = self._prec_code()
        a.mode = _lib.MODE_STOCK
        a.B = B
        a.weights = _lib.ptr(arena)
        a.packed = _lib.ptr(packed)
        a.workspace = _lib.ptr(ws)
        a.workspace_bytes = ws.numel()
        own = self._kv_owner
        if own is not None:
            a.k_cache = _lib.ptr(own._k_store)
            a.v_cache = _lib.ptr(own._v_store)
        return a

    # -- checkpoints ------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, precision: Optional[str] = None,
                             **_lightning_kwargs) -> "ScoreModule":
        """A reference (PyTorch-Lightning) checkpoint -> ScoreModule, like ``LightningModule.load_from_checkpoint``
        as cmd/sample.py:71-74 calls it: the constructor arguments come from the checkpoint's
        ``hyper_parameters`` (score_models.py:72 ``save_hyperparameters``: shape, d_model / num_layers / n_head,
        ``fourier_noise_scaling`` and the pickled noise scheduler -- VP or VE with its own constants), the weights
        from its ``state_dict``.  Lightning / omegaconf / the reference package need not be installed: their
        classes are unpickled as plain containers and ``fdiff.schedulers.sde.*`` as this package's schedulers."""
        from .checkpoint import load_lightning_checkpoint

        ckpt = load_lightning_checkpoint(checkpoint_path)
        hp = dict(ckpt.get("hyper_parameters") or {})
        missing = [k for k in ("n_channels", "max_len", "noise_scheduler") if k not in hp]
        if missing:
            raise _lib.FFDError(f"{checkpoint_path}: hyper_parameters lack {missing}; not a ScoreModule checkpoint")
        sched = hp["noise_scheduler"]
        if not isinstance(sched, SDE):
            raise NotImplementedError(f"Scheduler {sched!r} not implemented yet, cannot set time encoder.")
        known = ("n_channels", "max_len", "fourier_noise_scaling", "d_model", "num_layers", "n_head",
                 "num_training_steps", "lr_max", "likelihood_weighting")
        kwargs = {k: hp[k] for k in known if k in hp}
        model = cls(noise_scheduler=sched, precision=precision, **kwargs)
        sd = {k: v for k, v in ckpt["state_dict"].items() if not k.startswith("cached_backbone")}
        model.load_state_dict(sd, strict=True)
        if getattr(sched, "G", None) is None:
            sched.set_noise_scaling(model.max_len)
        return model

    # -- the reference surface --------------------------------------------------------
    def forward(self, batch: DiffusableBatch, recompute_tokens: Optional[Set[int]] = None, step: int = 0,
                return_crf: bool = False, t: Optional[float] = None
                ) -> Union[torch.Tensor, Tuple[torch.Tensor, Optional[torch.Tensor]]]:
        """score_models.py:79-119.  One timestep per batch (the sampler's contract, sampler.py:59-60); ``t`` (extra,
        optional) gives it as a host float, which skips the device-side equality check and its synchronisation."""
        X = batch.X
        assert X.size()[1:] == (self.max_len, self.n_channels), (
            f"X has wrong shape, should be {(X.size(0), self.max_len, self.n_channels)}, but is {X.size()}")
        timesteps = batch.timesteps
        assert timesteps is not None and timesteps.size(0) == len(batch)
        dev = self.device
        if dev.type != "cuda":
            raise _lib.FFDError("ScoreModule.forward needs the model on a CUDA device (no CPU fallback)")
        X = X.to(device=dev, dtype=torch.float32).contiguous()
        B = X.shape[0]
        if self.use_cache and recompute_tokens is not None and self.cache is not None \
                and getattr(self.cache, "event_driven", False):
            raise NotImplementedError(
                "event_driven=True decides the recompute set on the device inside the fused step loop "
                "(DiffusionSampler.sample / run_steps -> ffd_sample_steps); the per-step forward would silently reuse "
                "stale sample-0 K/V rows, so it is refused")
        if t is not None:
            te = self.time_embedding(torch.tensor([t], dtype=torch.float32))
        else:
            if B > 0 and not bool(torch.min(timesteps) == torch.max(timesteps)):
                raise _lib.FFDError("the CUDA score model evaluates one shared timestep per batch (sampler.py:59-60)")
            te = self.time_embedding(timesteps[:1]) if B > 0 else torch.zeros(1, self.d_model, device=dev)
        a = self.forward_args(B)
        a.X = _lib.ptr(X)
        a.time_emb = _lib.ptr(te)
        score = torch.empty_like(X)
        a.score = _lib.ptr(score)
        crf = None
        cached = self.use_cache and recompute_tokens is not None and self.cached_backbone is not None
        mask = None
        if cached:
            own = self._kv_owner
            S = sorted(set(int(i) for i in recompute_tokens))
            assert all(0 <= i < self.max_len for i in S)
            a.mode = _lib.MODE_CACHED
            a.n_own = len(S)
            if 0 < len(S) < self.max_len and S != list(range(len(S))):
                m = torch.zeros(self.max_len, dtype=torch.int32)
                m[S] = 1
                mask = m.to(dev)
                a.own_mask = _lib.ptr(mask)
            crf = torch.empty(len(self.backbone.layers), self.max_len, self.d_model, device=dev)
            a.crf_out = _lib.ptr(crf)
        _lib.check(_lib.load().ffd_score_forward(a, _lib.stream_handle(dev)))
        if cached:
            own.account_forward(len(S), S)
        if return_crf:
            return score, crf
        return score

    def enable_caching(self, cache: Optional[E2CRFCache] = None, **cache_kwargs) -> None:
        """score_models.py:202-283.  A new cache object is created on every call, but -- as in
        the reference, whose cached layers keep the cache they were built with -- K/V rows and
        the hit/recompute counters stay with the FIRST cache created for this model."""
        if cache is None:
            cache = E2CRFCache(max_len=self.max_len, num_layers=len(self.backbone.layers), device=self.device,
                               **cache_kwargs)
        self.cache = cache
        self.use_cache = True
        if self.cached_backbone is None:
            self.cached_backbone = _CachedBackboneMarker(len(self.backbone.layers))
            self._kv_owner = cache
        d = self.dims
        for c in {id(cache): cache, id(self._kv_owner): self._kv_owner}.values():
            if c._k_store is None or c._k_store.device != self.device:
                c.device = self.device
                c.attach_stores(d.n_head, d.d_model // d.n_head, d.d_model)

    def disable_caching(self) -> None:
        self.use_cache = False
        self.cache = None
