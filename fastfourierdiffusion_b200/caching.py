"""E2-CRF cache -- host mirror of ``fdiff.utils.caching.E2CRFCache``.

What the reference's cache does on the sampling path is:
a recompute schedule that is a pure function of the step index
(caching.py:131-181), a sample-0 K/V store of shape (num_layers, nhead,
max_len, head_dim) (caching.py:251-396), a CRF store written when
``current_step % R == 0`` (caching.py:459-480) and counters
(caching.py:599-636).  The K/V and CRF tensors live on the device and are
written by the kernels; the schedule and the counters are deterministic, so
they stay on the host and the step loop never synchronises.

FreqCa code of the reference is never reached from the sampler (SURVEY 2
row 5) and is not mirrored; the corresponding constructor arguments are
accepted and stored so that ``cache_kwargs`` keep working.  The event
intensity (caching.py:524-559; not called by the reference's sampler either)
is provided as a device kernel: ``compute_event_intensity`` mirrors the
reference call, ``event_intensity_device`` keeps the value on the device.

``event_driven=True`` (not a reference argument; default off) switches the
schedule to the one the reference's PAPER describes (section 3.3, Eq. 14-18,
Algorithm 1) and decides it on the device: per-token change of the final-layer
CRF against energy-weighted thresholds, an error-feedback accumulator, random
probes and the periodic gamma-correction -- ``ffd_event_decide`` /
``ffd_event_refresh_kv`` (include/ffd_b200.h).  The host never reads a
decision; ``get_event_stats`` is the one optional read-back, after sampling.
"""

from __future__ import annotations

from typing import Optional, Set

import torch


class E2CRFCache:
    def __init__(self, num_layers: int, max_len: int, device: torch.device, K: int = 5, R: int = 10,
                 tau_0: float = 0.1, tau_warn: float = 0.5, use_freqca: bool = False, freq_decomp: str = "dct",
                 low_freq_ratio: float = 0.3, max_history: int = 10, hermite_order: int = 3,
                 freq_decomp_interval: int = 10, use_fresca_in_cache: bool = False,
                 fresca_adaptive_threshold: bool = False, event_driven: bool = False,
                 event_probe_ratio: float = 0.0, event_accumulate: bool = True, event_eps: float = 1e-6,
                 event_eta: float = 1e-6, event_seed: int = 0):
        if use_freqca:
            raise NotImplementedError("FreqCa is off the sampling path of the reference and is not built")
        self.num_layers = num_layers
        self.max_len = max_len
        self.device = device
        self.K = K
        self.R = R
        self.tau_0 = tau_0
        self.tau_warn = tau_warn
        self.use_freqca = use_freqca
        self.freq_decomp = freq_decomp
        self.low_freq_ratio = low_freq_ratio
        self.max_history = max_history
        self.hermite_order = hermite_order
        self.freq_decomp_interval = freq_decomp_interval
        self.use_fresca_in_cache = use_fresca_in_cache
        self.fresca_adaptive_threshold = fresca_adaptive_threshold
        if event_driven and R <= 0:
            raise ValueError("event_driven needs a positive error-feedback period R")
        self.event_driven = bool(event_driven)
        self.event_probe_ratio = float(event_probe_ratio)
        self.event_accumulate = bool(event_accumulate)
        self.event_eps = float(event_eps)
        self.event_eta = float(event_eta)
        self.event_seed = int(event_seed)
        self._event_state: Optional[dict] = None  # device buffers of ffd_event_args
        # device stores, attached by ScoreModule.enable_caching (shape needs nhead/head_dim)
        self._k_store: Optional[torch.Tensor] = None  # [NL, H, L, hd], zero = never stored
        self._v_store: Optional[torch.Tensor] = None
        self._crf_store: Optional[torch.Tensor] = None  # [NL, L, d]
        self._populated = False  # reference: k_cache_tensor is None until the first store
        self._crf_valid = False
        self._valid = torch.zeros(num_layers, max_len, dtype=torch.bool)  # host copy of cache_valid
        self.stats = {"recompute_count": 0, "cache_hit_count": 0}
        self.current_step = 0

    # -- device stores ---------------------------------------------------------
    def attach_stores(self, n_head: int, head_dim: int, d_model: int) -> None:
        dev = self.device
        self._k_store = torch.zeros(self.num_layers, n_head, self.max_len, head_dim, device=dev)
        self._v_store = torch.zeros_like(self._k_store)
        self._crf_store = torch.zeros(self.num_layers, self.max_len, d_model, device=dev)

    def event_args(self, d_model: int):
        """ffd_event_args over this cache's device buffers (allocated on first use)."""
        from . import _lib

        dev = self.device
        st = self._event_state
        if st is None or st["z_prev"].device != dev or st["z_prev"].shape[1] != d_model:
            st = {"z_prev": torch.zeros(self.max_len, d_model, device=dev),
                  "accum": torch.zeros(self.max_len, device=dev),
                  "weight": torch.zeros(self.max_len, device=dev),
                  "stats": torch.zeros(_lib.EVENT_STATS, device=dev)}
            self._event_state = st
        e = _lib.EventArgs()
        e.z_prev, e.accum, e.weight, e.stats = (_lib.ptr(st[k]) for k in ("z_prev", "accum", "weight", "stats"))
        e.K, e.R, e.accumulate = int(self.K), int(self.R), int(self.event_accumulate)
        e.tau_0, e.tau_warn, e.eps, e.eta = float(self.tau_0), float(self.tau_warn), self.event_eps, self.event_eta
        e.probe_ratio, e.seed = self.event_probe_ratio, self.event_seed
        return e

    def get_event_stats(self) -> dict:
        """Counters the decision kernel keeps on the device (one read-back; never needed by the step loop)."""
        if self._event_state is None:
            return {}
        s = self._event_state["stats"].tolist()
        n = max(s[0], 1.0)
        return {"steps_decided": int(s[0]), "tokens_recomputed": int(s[1]), "tokens_corrected": int(s[2]),
                "feedback_steps": int(s[3]), "last_event_intensity": s[4], "mean_event_intensity": s[5] / n,
                "recompute_fraction": s[1] / (n * self.max_len)}

    @property
    def k_cache_tensor(self) -> Optional[torch.Tensor]:
        return self._k_store if self._populated else None

    @property
    def v_cache_tensor(self) -> Optional[torch.Tensor]:
        return self._v_store if self._populated else None

    @property
    def cache_valid(self) -> Optional[torch.Tensor]:
        return self._valid.to(self.device) if self._populated else None

    @property
    def crf_cache(self) -> Optional[torch.Tensor]:
        return self._crf_store if self._crf_valid else None

    def reset(self) -> None:
        """caching.py:115-129 (tensors dropped there; zeroed here, which reads the same)."""
        if self._k_store is not None:
            self._k_store.zero_()
            self._v_store.zero_()
        self._populated = False
        self._crf_valid = False
        self._valid.zero_()
        self.stats = {"recompute_count": 0, "cache_hit_count": 0}
        self.current_step = 0

    # -- schedule ------------------------------------------------------------------
    def determine_recompute_set(self, x_tilde: Optional[torch.Tensor], event_intensity: float,
                                step: int) -> Set[int]:
        """caching.py:131-181: all tokens at step 0, the first min(2K, L) tokens every
        ``500 if R < 100 else R`` steps, nothing otherwise.  ``x_tilde`` and
        ``event_intensity`` are ignored, as in the reference."""
        if step == 0:
            return set(range(self.max_len))
        if self.event_driven:
            return set()  # which tokens are refreshed is decided on the device (ffd_event_decide)
        interval = 500 if self.R < 100 else self.R
        if step % interval == 0:
            return set(range(min(2 * min(self.K, self.max_len), self.max_len)))
        return set()

    def crf_stored_at(self, step: int) -> bool:
        """update_crf keeps the CRF iff current_step % R == 0 (caching.py:474); R == 0
        raises ZeroDivisionError exactly like the reference."""
        return step % self.R == 0

    # -- host-side bookkeeping of what the kernels did --------------------------
    def account_forward(self, n_own: int, own_idx=None) -> None:
        """Counter / validity updates of one cached-mode forward over all layers
        (caching.py:283,299,389-396; branch rules cached_transformer.py:142-305)."""
        L, NL = self.max_len, self.num_layers
        if n_own == L:
            self.stats["recompute_count"] += NL * L
            self._valid[:] = True
            self._populated = True
        elif n_own > 0.8 * L:
            return
        elif n_own == 0:
            if self._populated:
                self.stats["cache_hit_count"] += NL * L
        else:
            for _ in range(NL):
                if self._populated and L - n_own > 0:
                    self.stats["cache_hit_count"] += L - n_own
                self.stats["recompute_count"] += n_own
                self._populated = True
            idx = list(range(n_own)) if own_idx is None else list(own_idx)
            self._valid[:, idx] = True

    def update_crf(self, crf: Optional[torch.Tensor], timestep: Optional[float] = None) -> None:
        """caching.py:459-480.  On the fused path the kernel has already written the
        CRF rows into the device store; a tensor passed here is copied in."""
        if self.crf_stored_at(self.current_step):
            if crf is not None and self._crf_store is not None and crf.data_ptr() != self._crf_store.data_ptr():
                self._crf_store.copy_(crf)
            self._crf_valid = True

    # -- event intensity ------------------------------------------
    def event_intensity_device(self, crf: torch.Tensor) -> torch.Tensor:
        """[avg_energy, intensity] of ``crf`` against the stored CRF as a 2-element DEVICE tensor (no host sync):
        mean over (layer, token) of ||crf - crf_cache||_2, and min(1, that / tau_0)  (caching.py:541-557)."""
        from . import _lib

        prev = self.crf_cache
        if prev is None:
            raise _lib.FFDError("event_intensity_device needs a stored CRF")
        dev = _lib.require_cuda()
        crf = crf.to(device=dev, dtype=torch.float32).contiguous()
        if crf.shape != prev.shape:
            raise ValueError(f"crf shape {tuple(crf.shape)} != stored {tuple(prev.shape)}")
        out = torch.empty(2, dtype=torch.float32, device=dev)
        d = crf.shape[-1]
        _lib.check(_lib.load().ffd_event_intensity(_lib.ptr(crf), _lib.ptr(prev.contiguous()), crf.numel() // d, d,
                                                   float(self.tau_0), _lib.ptr(out), _lib.stream_handle()))
        return out

    def compute_event_intensity(self, crf: torch.Tensor, step: int) -> float:
        """caching.py:524-559: 1.0 at step 0 / 0.1 later while no CRF is stored, else min(1, mean row energy of
        the CRF change / tau_0).  Returns a host float like the reference (one sync)."""
        if self.crf_cache is None:
            return 0.1 if step > 0 else 1.0
        return float(self.event_intensity_device(crf)[1].item())

    def get_cache_stats(self) -> dict:
        """caching.py:599-636."""
        total = self.stats["recompute_count"] + self.stats["cache_hit_count"]
        ratio = float(self._valid.float().mean().item()) if self._populated else 0.0
        if ratio >= 1.0:
            ratio = 0.99  # the reference caps a fully valid cache at 99 % (caching.py:623-628)
        return {
            "cache_hit_ratio": self.stats["cache_hit_count"] / total if total > 0 else 0.0,
            "cache_ratio": ratio,
            "recompute_count": self.stats["recompute_count"],
            "cache_hit_count": self.stats["cache_hit_count"],
            "current_step": self.current_step,
        }
