"""CPU oracle for the FastFourierDiffusion reverse-SDE sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``fastfourierdiffusion_b200/`` may import
this module; it is imported by ``tests/``, by ``__graft_entry__.smoke()`` and by
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there
only as the checker or as the timed CPU baseline, never as the product path.

It is a plain-numpy restatement (no torch ops on the arithmetic path) of the
reference's algorithm.  Every function cites the reference file:line it
follows (paths relative to the upstream repo root).  Parity status: PINNED --
``oracle/make_golden.py`` runs the unmodified reference classes (imported from
the upstream checkout with two stub modules) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.

The arithmetic dtype is a parameter (``np.float32`` mirrors the reference,
``np.float64`` separates our error from the reference's own fp32 rounding).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

# --------------------------------------------------------------------------- #
# Fourier boundary transforms  (src/fdiff/utils/fourier.py:8-94)
# --------------------------------------------------------------------------- #


def n_real(max_len: int) -> int:
    """Number of real components kept by the packed DFT (fourier.py:66)."""
    return int(math.ceil((max_len + 1) / 2))


def dft_matrix(max_len: int, dtype=np.float64) -> np.ndarray:
    """U such that ``dft(x)[b,:,c] = U @ x[b,:,c]``.

    fourier.py:24-46: ``rfft(norm="ortho")``, real parts of bins 0..n_r-1, then
    imaginary parts of bins 1.. (the last one dropped when L is even).
    """
    L = max_len
    nr = n_real(L)
    n = np.arange(L, dtype=np.float64)
    U = np.zeros((L, L), dtype=np.float64)
    for k in range(nr):
        U[k] = np.cos(2.0 * np.pi * k * n / L) / math.sqrt(L)
    for j in range(1, L - nr + 1):
        U[nr - 1 + j] = -np.sin(2.0 * np.pi * j * n / L) / math.sqrt(L)
    return U.astype(dtype)


def idft_matrix(max_len: int, dtype=np.float64) -> np.ndarray:
    """V such that ``idft(x)[b,:,c] = V @ x[b,:,c]``  (fourier.py:55-94).

    ``irfft(n=L, norm="ortho")`` of the re-assembled half spectrum equals
    ``(diag(w) U)^T`` with w = 2 except the DC bin (and Nyquist if L is even).
    """
    L = max_len
    nr = n_real(L)
    U = dft_matrix(L, np.float64)
    w = np.full(L, 2.0)
    w[0] = 1.0
    if L % 2 == 0:
        w[L // 2] = 1.0  # Nyquist real part lives at row L/2 (= nr-1)
    return (U * w[:, None]).T.astype(dtype)


def dft(x: np.ndarray) -> np.ndarray:
    """fourier.py:8-52 on ``x[B, L, C]``."""
    U = dft_matrix(x.shape[1], x.dtype)
    return np.einsum("kn,bnc->bkc", U, x)


def idft(x: np.ndarray) -> np.ndarray:
    """fourier.py:55-94 on ``x[B, L, C]``."""
    V = idft_matrix(x.shape[1], x.dtype)
    return np.einsum("nk,bkc->bnc", V, x)


# --------------------------------------------------------------------------- #
# SDE schedule  (src/fdiff/schedulers/sde.py)
# --------------------------------------------------------------------------- #


def noise_scaling(max_len: int, fourier_noise_scaling: bool) -> np.ndarray:
    """G of sde.py:42-60 as float32 (the reference builds it in fp32)."""
    G = np.ones(max_len, dtype=np.float32)
    if fourier_noise_scaling:
        G = (np.float32(1.0 / math.sqrt(2.0)) * G).astype(np.float32)
        G[0] = np.float32(G[0] * np.float32(math.sqrt(2.0)))
        if max_len % 2 == 0:
            G[max_len // 2] = np.float32(G[max_len // 2] * np.float32(math.sqrt(2.0)))
    return G


def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """``torch.linspace(start, end, steps)`` in fp32 (sde.py:63).

    ATen's CPU kernel computes ``step=(end-start)/(steps-1)`` in fp32 and fills
    the first half from ``start`` upwards and the second half from ``end``
    downwards.  Its vectorised inner loop adds ``step*lane`` to a per-chunk
    base, so single entries can differ from this scalar restatement by one ulp
    depending on the host's SIMD width; callers that need the reference's exact
    grid pass ``ts`` (the golden fixtures store it).
    """
    out = np.empty(steps, dtype=np.float32)
    if steps == 1:
        out[0] = np.float32(start)
        return out
    s = np.float32(start)
    e = np.float32(end)
    step = np.float32((e - s) / np.float32(steps - 1))
    half = steps // 2
    for i in range(steps):
        if i < half:
            out[i] = np.float32(s + np.float32(step * np.float32(i)))
        else:
            out[i] = np.float32(e - np.float32(step * np.float32(steps - 1 - i)))
    return out


@dataclass
class SDEConfig:
    """VP (sde.py:168-186) or VE (sde.py:90-107) scheduler parameters."""

    kind: str = "vp"  # "vp" | "ve"
    beta_min: float = 0.1
    beta_max: float = 20.0
    sigma_min: float = 0.01
    sigma_max: float = 50.0
    fourier_noise_scaling: bool = False
    eps: float = 1e-5


def sde_step_coeffs(cfg: SDEConfig, t: float) -> Tuple[float, float]:
    """Per-step python-float scalars.

    VP: (beta, sqrt(beta))              sde.py:212-213,228-230
    VE: (0, sqrt_derivative)            sde.py:143-147
    """
    if cfg.kind == "vp":
        beta = cfg.beta_min + t * (cfg.beta_max - cfg.beta_min)
        return beta, math.sqrt(beta)
    sd = cfg.sigma_min * math.sqrt(2 * math.log(cfg.sigma_max / cfg.sigma_min)) * (
        cfg.sigma_max / cfg.sigma_min
    ) ** t
    return 0.0, sd


def sde_step(cfg: SDEConfig, G: np.ndarray, score: np.ndarray, t: float, x: np.ndarray,
             z: np.ndarray, step_size: np.float32, dtype=np.float32) -> np.ndarray:
    """Euler-Maruyama reverse step, same op order as the reference.

    VP sde.py:228-244, VE sde.py:143-163.  The reference multiplies by dense
    diagonal matrices; a diagonal matmul is one product plus exact zeros, so it
    is restated element-wise.  With ``dtype=np.float32`` every intermediate is
    rounded where the reference rounds, so the result is bit-identical.
    """
    f = dtype
    beta, coef = sde_step_coeffs(cfg, t)
    # diag_embed(coef * G): python float * fp32 tensor -> fp32
    diff = (f(coef) * G.astype(f)).astype(f)[None, :, None]
    diff2 = (diff * diff).astype(f)
    x = x.astype(f)
    score = score.astype(f)
    z = z.astype(f)
    if cfg.kind == "vp":
        drift = ((f(-0.5 * beta) * x).astype(f) - (diff2 * score).astype(f)).astype(f)
    else:
        drift = (-(diff2 * score).astype(f)).astype(f)
    dt = f(step_size)
    out = (x - (drift * dt).astype(f)).astype(f)
    out = (out + (f(np.sqrt(dt)) * (diff * z).astype(f)).astype(f)).astype(f)
    return out


def prior_sample(cfg: SDEConfig, G: np.ndarray, z: np.ndarray, dtype=np.float32) -> np.ndarray:
    """sde.py:79-87 (``G_matrix @ z``) and the VE rescale sde.py:125-127."""
    x = (G.astype(dtype)[None, :, None] * z.astype(dtype)).astype(dtype)
    if cfg.kind == "ve":
        x = (dtype(cfg.sigma_max) * x).astype(dtype)
    return x


# --------------------------------------------------------------------------- #
# Score model pieces  (src/fdiff/models/transformer.py, score_models.py)
# --------------------------------------------------------------------------- #


def posemb_renorm(table: np.ndarray, max_norm: float, iters: int = 1) -> np.ndarray:
    """``nn.Embedding(max_norm=sqrt(d))`` in-place renorm (transformer.py:13-15,27).

    ATen ``embedding_renorm_``: rows whose 2-norm exceeds ``max_norm`` are
    scaled by ``max_norm / (norm + 1e-7)``.  The reference re-applies it on
    every forward; ``iters`` emulates that.
    """
    t = table.copy()
    for _ in range(iters):
        norm = np.sqrt((t.astype(np.float64) ** 2).sum(axis=1)).astype(t.dtype)
        scale = np.where(norm > max_norm, t.dtype.type(max_norm) / (norm + t.dtype.type(1e-7)), 1.0)
        t = (t * scale[:, None].astype(t.dtype)).astype(t.dtype)
    return t


def time_embedding(t: float, W: np.ndarray, dense_w: np.ndarray, dense_b: np.ndarray,
                   d_model: int, dtype=np.float32) -> np.ndarray:
    """GaussianFourierProjection (transformer.py:77-91) for one shared t -> [d]."""
    f = dtype
    # ((t * W) * 2) * pi, each product rounded in the tensor dtype
    proj = (f(np.float32(t)) * W.astype(f)).astype(f)
    proj = (proj * f(2)).astype(f)
    proj = (proj * f(np.pi)).astype(f)
    emb = np.concatenate([np.sin(proj), np.cos(proj)]).astype(f)[:d_model]
    return (dense_w.astype(f) @ emb + dense_b.astype(f)).astype(f)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * w + b


def softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


@dataclass
class LayerWeights:
    in_proj_w: np.ndarray  # [3d, d]
    in_proj_b: np.ndarray  # [3d]
    out_w: np.ndarray  # [d, d]
    out_b: np.ndarray
    w1: np.ndarray  # [F, d]
    b1: np.ndarray
    w2: np.ndarray  # [d, F]
    b2: np.ndarray
    n1w: np.ndarray
    n1b: np.ndarray
    n2w: np.ndarray
    n2b: np.ndarray


@dataclass
class ModelWeights:
    pos: np.ndarray  # [L, d]  (before renorm)
    time_W: np.ndarray  # [ceil(d/2)]
    time_dense_w: np.ndarray
    time_dense_b: np.ndarray
    emb_w: np.ndarray  # [d, C]
    emb_b: np.ndarray
    unemb_w: np.ndarray  # [C, d]
    unemb_b: np.ndarray
    layers: List[LayerWeights]
    n_head: int

    @property
    def d_model(self) -> int:
        return self.emb_w.shape[0]

    @staticmethod
    def from_state_dict(sd: Dict[str, np.ndarray], n_head: int, dtype=np.float32) -> "ModelWeights":
        """Unpack the 128-key ``ScoreModule.state_dict()`` (score_models.py:55-66)."""
        g = lambda k: np.asarray(sd[k], dtype=dtype)
        layers = []
        i = 0
        while f"backbone.layers.{i}.linear1.weight" in sd:
            p = f"backbone.layers.{i}."
            layers.append(LayerWeights(
                g(p + "self_attn.in_proj_weight"), g(p + "self_attn.in_proj_bias"),
                g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"),
                g(p + "linear1.weight"), g(p + "linear1.bias"),
                g(p + "linear2.weight"), g(p + "linear2.bias"),
                g(p + "norm1.weight"), g(p + "norm1.bias"),
                g(p + "norm2.weight"), g(p + "norm2.bias")))
            i += 1
        return ModelWeights(
            g("pos_encoder.embedding.weight"), g("time_encoder.W"),
            g("time_encoder.dense.weight"), g("time_encoder.dense.bias"),
            g("embedder.weight"), g("embedder.bias"),
            g("unembedder.weight"), g("unembedder.bias"), layers, n_head)


def _split_heads(x: np.ndarray, H: int) -> np.ndarray:
    B, L, d = x.shape
    return x.reshape(B, L, H, d // H).transpose(0, 2, 1, 3)  # [B,H,L,hd]


def _merge_heads(x: np.ndarray) -> np.ndarray:
    B, H, L, hd = x.shape
    return x.transpose(0, 2, 1, 3).reshape(B, L, H * hd)


def _attn(q, k, v):
    hd = q.shape[-1]
    s = np.einsum("bhid,bhjd->bhij", q, k) / q.dtype.type(math.sqrt(hd))
    return np.einsum("bhij,bhjd->bhid", softmax(s), v)


def _ffn_ln(lw: LayerWeights, x: np.ndarray) -> np.ndarray:
    hid = np.maximum(x @ lw.w1.T + lw.b1, 0)
    return layer_norm(x + hid @ lw.w2.T + lw.b2, lw.n2w, lw.n2b)


def stock_layer(lw: LayerWeights, x: np.ndarray, H: int) -> np.ndarray:
    """Post-norm ``nn.TransformerEncoderLayer`` in eval mode
    (score_models.py:61-66; cached_transformer.py:123-134)."""
    d = x.shape[-1]
    qkv = x @ lw.in_proj_w.T + lw.in_proj_b
    q, k, v = (_split_heads(qkv[..., i * d:(i + 1) * d], H) for i in range(3))
    a = _merge_heads(_attn(q, k, v)) @ lw.out_w.T + lw.out_b
    x = layer_norm(x + a, lw.n1w, lw.n1b)
    return _ffn_ln(lw, x)


class KVStore:
    """Tensor KV cache of E2CRFCache (caching.py:87-92,251-300,302-396)."""

    def __init__(self, num_layers: int, max_len: int):
        self.num_layers = num_layers
        self.max_len = max_len
        self.reset()

    def reset(self) -> None:  # caching.py:115-129
        self.k: Optional[np.ndarray] = None  # [NL,H,L,hd]
        self.v: Optional[np.ndarray] = None
        self.valid: Optional[np.ndarray] = None
        self.recompute_count = 0
        self.cache_hit_count = 0

    def set_batch(self, layer: int, idx: Sequence[int], k: np.ndarray, v: np.ndarray) -> None:
        """Keep sample 0 only (caching.py:326-328), lazy zero alloc (:364-377)."""
        k0, v0 = k[0], v[0]  # [H, n, hd]
        if self.k is None:
            H, _, hd = k0.shape
            self.k = np.zeros((self.num_layers, H, self.max_len, hd), dtype=k0.dtype)
            self.v = np.zeros_like(self.k)
            self.valid = np.zeros((self.num_layers, self.max_len), dtype=bool)
        idx = list(idx)
        self.k[layer][:, idx, :] = k0
        self.v[layer][:, idx, :] = v0
        self.valid[layer, idx] = True
        self.recompute_count += len(idx)  # caching.py:396

    def get(self, layer: int, idx: Sequence[int]):
        if self.k is None:
            return None
        idx = list(idx)
        self.cache_hit_count += len(idx)  # caching.py:283,299
        return self.k[layer][:, idx, :], self.v[layer][:, idx, :]


def cached_layer(lw: LayerWeights, x: np.ndarray, H: int, S: Sequence[int], store: KVStore,
                 layer: int) -> np.ndarray:
    """CachedTransformerEncoderLayer.forward (cached_transformer.py:106-329)."""
    B, L, d = x.shape
    S = sorted(S)
    nS = len(S)
    wq, wk, wv = (lw.in_proj_w[i * d:(i + 1) * d] for i in range(3))
    bq, bk, bv = (lw.in_proj_b[i * d:(i + 1) * d] for i in range(3))
    if nS == L:
        # :142-191 -- stock layer, then K/V projected from the layer OUTPUT
        y = stock_layer(lw, x, H)
        store.set_batch(layer, range(L), _split_heads(y @ wk.T + bk, H), _split_heads(y @ wv.T + bv, H))
        return y
    if nS > 0.8 * L:
        # :196-220 -- stock layer, cache untouched
        return stock_layer(lw, x, H)
    q = _split_heads(x @ wq.T + bq, H)
    hd = d // H
    if nS == 0:
        # :237-258 pure cache: sample-0 K/V broadcast over the batch
        got = store.get(layer, range(L))
        if got is None:
            kf = np.zeros((B, H, L, hd), dtype=x.dtype)
            vf = np.zeros_like(kf)
        else:
            kf = np.broadcast_to(got[0][None], (B, H, L, hd))
            vf = np.broadcast_to(got[1][None], (B, H, L, hd))
    else:
        # :259-305 mixed: zeros, cached rows for L\S, recomputed rows for S from the layer INPUT
        kf = np.zeros((B, H, L, hd), dtype=x.dtype)
        vf = np.zeros_like(kf)
        rest = [i for i in range(L) if i not in set(S)]
        if rest:
            got = store.get(layer, rest)
            if got is not None:
                kf[:, :, rest, :] = got[0][None]
                vf[:, :, rest, :] = got[1][None]
        xs = x[:, S, :]
        kr = _split_heads(xs @ wk.T + bk, H)
        vr = _split_heads(xs @ wv.T + bv, H)
        kf[:, :, S, :] = kr
        vf[:, :, S, :] = vr
        store.set_batch(layer, S, kr, vr)
    a = _merge_heads(_attn(q, kf, vf)) @ lw.out_w.T + lw.out_b  # :309-317
    x = layer_norm(x + a, lw.n1w, lw.n1b)  # :320-321
    return _ffn_ln(lw, x)  # :325-327


# --------------------------------------------------------------------------- #
# E2-CRF schedule / stats  (src/fdiff/utils/caching.py)
# --------------------------------------------------------------------------- #


def determine_recompute_set(step: int, K: int, R: int, max_len: int) -> List[int]:
    """caching.py:131-181 -- a pure function of the step index."""
    if step == 0:
        return list(range(max_len))
    interval = 500 if R < 100 else R
    if step % interval == 0:
        return list(range(min(2 * min(K, max_len), max_len)))
    return []


def cache_stats(store: KVStore, current_step: int) -> Dict[str, float]:
    """caching.py:599-636 (FreqCa keys omitted: off on the sampling path)."""
    total = store.recompute_count + store.cache_hit_count
    ratio = float(store.valid.astype(np.float32).mean()) if store.valid is not None else 0.0
    if ratio >= 1.0:
        ratio = 0.99
    return {
        "cache_hit_ratio": store.cache_hit_count / total if total > 0 else 0.0,
        "cache_ratio": ratio,
        "recompute_count": store.recompute_count,
        "cache_hit_count": store.cache_hit_count,
        "current_step": current_step,
    }


# --------------------------------------------------------------------------- #
# Score model + sampler  (score_models.py:79-194, sampler.py:48-228)
# --------------------------------------------------------------------------- #


# --------------------------------------------------------------------------- #
# Event-driven schedule of the reference's PAPER (submission/TSD_Cache_Paper_Final.pdf
# section 3.3: Eq. 14 intensity, Eq. 15 recompute set, Eq. 17-18 error feedback,
# Algorithm 1).  The reference's code never runs it (caching.py:131-181 decides by
# step index; caching.py:524-559 is dead), so this part of the oracle restates the
# paper's equations, not reference code: parity for it is UNPINNED by reference runs
# and is anchored on the closed forms in tests/test_event_cpu.py instead.
# --------------------------------------------------------------------------- #


@dataclass
class EventConfig:
    K: int = 5
    R: int = 10
    tau_0: float = 0.1
    tau_warn: float = 0.5
    eps: float = 1e-6
    eta: float = 1e-6
    probe_ratio: float = 0.0
    accumulate: bool = True
    seed: int = 0


class EventState:
    def __init__(self, max_len: int, d_model: int, dtype=np.float32):
        self.z_prev = np.zeros((max_len, d_model), dtype=dtype)
        self.accum = np.zeros(max_len, dtype=dtype)
        self.weight = np.zeros(max_len, dtype=dtype)
        self.stats = np.zeros(8, dtype=np.float64)
        self.last: Dict[str, np.ndarray] = {}


def event_uniform(seed: int, step: int, max_len: int) -> np.ndarray:
    """RandomProbe draw per token: Philox(counter=(k, 0, step, 0), key=seed), first word -> (0, 1) in fp32."""
    k = np.arange(max_len, dtype=np.uint32)
    ctr = np.stack([k, np.zeros_like(k), np.full_like(k, step), np.zeros_like(k)], axis=1)
    key = np.stack([np.full(max_len, seed & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(max_len, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=1)
    r = philox4x32_10(ctr, key)[:, 0]
    return (r.astype(np.float32) + np.float32(0.5)) * np.float32(2.3283064365386963e-10)


def event_decide(cfg: EventConfig, st: EventState, z: np.ndarray, x0: np.ndarray, step: int, first: bool) -> None:
    """Decision after the score model of step `step`: z = final-layer CRF of sample 0 [L, d], x0 = sample 0's
    state [L, C].  Leaves the next step's blend weights in st.weight."""
    dt = st.z_prev.dtype
    z = z.astype(dt)
    L = z.shape[0]
    if first:
        st.z_prev = z.copy()
        st.accum[:] = 0
        st.weight[:] = 0
        st.stats[:] = 0
        st.last = {}
        return
    diff = z - st.z_prev
    delta = np.sqrt((diff * diff).sum(axis=1))                       # ||z_i[k] - z_{i-1}[k]||_2
    r = (diff * diff).sum() / ((st.z_prev * st.z_prev).sum() + dt.type(cfg.eta))   # Eq. 14
    feedback = ((step + 1) % cfg.R == 0) or (r > cfg.tau_warn)       # Algorithm 1, line 16
    gamma = min(0.1, 0.5 * float(r))
    energy = (x0.astype(dt) ** 2).sum(axis=1)
    tau = dt.type(cfg.tau_0) / (dt.type(cfg.eps) + energy)           # Eq. 15 thresholds
    a = (st.accum if cfg.accumulate else 0) + delta
    k = np.arange(L)
    probe = np.zeros(L, dtype=bool)
    if cfg.probe_ratio > 0:
        probe = (k > cfg.K) & (event_uniform(cfg.seed, step, L) < np.float32(cfg.probe_ratio))
    rec = (k <= cfg.K) | (a > tau) | probe                           # Eq. 15
    st.weight = np.where(rec, 1.0, gamma if feedback else 0.0).astype(dt)
    st.accum = np.where(rec, 0.0, a).astype(dt)
    st.z_prev = z.copy()
    st.stats[0] += 1
    st.stats[1] += int(rec.sum())
    st.stats[2] += int(((~rec) & (st.weight > 0)).sum())
    st.stats[3] += 1 if feedback else 0
    st.stats[4] = float(r)
    st.stats[5] += float(r)
    st.last = {"r": np.asarray(r), "a": a, "tau": tau, "rec": rec, "probe": probe, "feedback": np.asarray(feedback),
               "gamma": np.asarray(gamma), "delta": delta}


def event_refresh_kv(lw: LayerWeights, h0: np.ndarray, weight: np.ndarray, store: KVStore, layer: int, H: int) -> None:
    """K/V rows of sample 0 for layer `layer` from that layer's input h0 [L, d] (the mixed branch's projection,
    cached_transformer.py:282-296), blended into the store: cache += weight * (new - cache)  (Eq. 18; weight 1 =
    recompute, 0 = reuse)."""
    L, d = h0.shape
    wk, wv = lw.in_proj_w[d:2 * d], lw.in_proj_w[2 * d:3 * d]
    bk, bv = lw.in_proj_b[d:2 * d], lw.in_proj_b[2 * d:3 * d]
    kn = _split_heads((h0 @ wk.T + bk)[None], H)[0]  # [H, L, hd]
    vn = _split_heads((h0 @ wv.T + bv)[None], H)[0]
    w = weight.astype(h0.dtype)[None, :, None]
    on = (weight > 0)[None, :, None]
    full = (weight >= 1)[None, :, None]
    store.k[layer] = np.where(full, kn, np.where(on, store.k[layer] + w * (kn - store.k[layer]), store.k[layer]))
    store.v[layer] = np.where(full, vn, np.where(on, store.v[layer] + w * (vn - store.v[layer]), store.v[layer]))


class OracleScoreModel:
    def __init__(self, w: ModelWeights, max_len: int, n_channels: int, dtype=np.float32,
                 renorm_iters: int = 3):
        self.w = w
        self.L = max_len
        self.C = n_channels
        self.dtype = dtype
        self.H = w.n_head
        d = w.d_model
        self.pos = posemb_renorm(w.pos.astype(np.float32), math.sqrt(d), renorm_iters).astype(dtype)

    def embed(self, X: np.ndarray, t: float) -> np.ndarray:
        """score_models.py:96-102."""
        w = self.w
        h = X.astype(self.dtype) @ w.emb_w.astype(self.dtype).T + w.emb_b.astype(self.dtype)
        h = h + self.pos[None]
        te = time_embedding(t, w.time_W, w.time_dense_w, w.time_dense_b, w.d_model, self.dtype)
        return h + te[None, None, :]

    def unembed(self, h: np.ndarray) -> np.ndarray:
        return h @ self.w.unemb_w.astype(self.dtype).T + self.w.unemb_b.astype(self.dtype)

    def forward(self, X: np.ndarray, t: float) -> np.ndarray:
        """No-cache path (score_models.py:109)."""
        h = self.embed(X, t)
        for lw in self.w.layers:
            h = stock_layer(_cast_layer(lw, self.dtype), h, self.H)
        return self.unembed(h)

    def forward_cached(self, X: np.ndarray, t: float, S: Sequence[int], store: KVStore):
        """_forward_with_cache (score_models.py:121-194): returns (score, crf[NL,L,d])."""
        h = self.embed(X, t)
        crf = []
        for i, lw in enumerate(self.w.layers):
            h = cached_layer(_cast_layer(lw, self.dtype), h, self.H, S, store, i)
            crf.append(h[0].copy())
        return self.unembed(h), np.stack(crf)


    def forward_event(self, X: np.ndarray, t: float, store: KVStore, weight: np.ndarray):
        """A pure-cache step of the event-driven schedule: every layer first refreshes the sample-0 K/V rows
        from its own input under `weight`, then runs the pure-cache branch.  Returns (score, z_L)."""
        h = self.embed(X, t)
        for i, lw in enumerate(self.w.layers):
            lw = _cast_layer(lw, self.dtype)
            event_refresh_kv(lw, h[0], weight, store, i, self.H)
            h = cached_layer(lw, h, self.H, [], store, i)
        return self.unembed(h), h[0].copy()


def _cast_layer(lw: LayerWeights, dtype) -> LayerWeights:
    if lw.w1.dtype == dtype:
        return lw
    return LayerWeights(*[getattr(lw, f).astype(dtype) for f in lw.__dataclass_fields__])


@dataclass
class OracleTrace:
    states: List[np.ndarray] = field(default_factory=list)  # X_t fed to step t
    scores: List[np.ndarray] = field(default_factory=list)
    recompute: List[List[int]] = field(default_factory=list)


class OracleSampler:
    """DiffusionSampler.sample (sampler.py:105-215) with injected noise."""

    def __init__(self, model: OracleScoreModel, sde: SDEConfig, use_cache: bool = False,
                 K: int = 5, R: int = 10, event: Optional[EventConfig] = None):
        self.event = event
        self.event_state = (EventState(model.L, model.w.d_model, np.dtype(model.dtype)) if event is not None
                            else None)
        self.model = model
        self.sde = sde
        self.use_cache = use_cache
        self.K, self.R = K, R
        self.G = noise_scaling(model.L, sde.fourier_noise_scaling)
        self.store = KVStore(len(model.w.layers), model.L)
        self.crf: Optional[np.ndarray] = None
        self.current_step = 0

    def step(self, X: np.ndarray, t: float, z: np.ndarray, step_size: np.float32,
             S: Optional[Sequence[int]], global_step: int) -> Tuple[np.ndarray, np.ndarray]:
        """reverse_diffusion_step (sampler.py:48-103)."""
        if self.use_cache and S is not None:
            self.current_step = global_step  # sampler.py:182
            if self.event is not None and len(S) == 0:
                score, zL = self.model.forward_event(X, t, self.store, self.event_state.weight)
                event_decide(self.event, self.event_state, zL, X[0], global_step, False)
                return sde_step(self.sde, self.G, score, t, X, z, step_size, self.model.dtype), score
            score, crf = self.model.forward_cached(X, t, S, self.store)
            if self.event is not None and len(S) == self.model.L:
                event_decide(self.event, self.event_state, crf[-1], X[0], global_step, True)
            if self.current_step % self.R == 0:  # update_crf, caching.py:474 (R=0 raises there too)
                self.crf = crf
        else:
            score = self.model.forward(X, t)
        Xn = sde_step(self.sde, self.G, score, t, X, z, step_size, self.model.dtype)
        return Xn, score

    def sample_batches(self, priors: Sequence[np.ndarray], noises: Sequence[Sequence[np.ndarray]],
                       n_steps: int, trace: Optional[OracleTrace] = None,
                       ts: Optional[np.ndarray] = None) -> np.ndarray:
        """One entry of ``priors``/``noises`` per sampler batch.  The cache is
        reset only for batch 0 and the global step keeps counting
        (sampler.py:151-153,210)."""
        if ts is None:
            ts = linspace_f32(1.0, self.sde.eps, n_steps)
        ts = np.asarray(ts, dtype=np.float32)
        step_size = np.float32(ts[0] - ts[1])  # sde.py:64
        outs = []
        gstep = 0
        for b, (z0, zs) in enumerate(zip(priors, noises)):
            X = prior_sample(self.sde, self.G, z0, self.model.dtype)
            if self.use_cache and b == 0:
                self.store.reset()
                self.crf = None
                gstep = 0
            for i in range(n_steps):
                t = float(ts[i])
                S = determine_recompute_set(gstep, self.K, self.R, self.model.L) if self.use_cache else None
                if trace is not None:
                    trace.states.append(X.copy())
                    trace.recompute.append(list(S) if S is not None else None)
                X, score = self.step(X, t, zs[i], step_size, S, gstep)
                if trace is not None:
                    trace.scores.append(score)
                gstep += 1
            outs.append(X)
        return np.concatenate(outs, axis=0)


# --------------------------------------------------------------------------- #
# FreSca score filter  (src/fdiff/utils/fresca.py)
# --------------------------------------------------------------------------- #


def fresca_cutoff(spec: np.ndarray, cutoff_ratio: float, strategy: str) -> int:
    """Largest low-band bin index Rc (fresca.py:38-60)."""
    n_freq = spec.shape[0]
    if strategy == "spatial":
        lim = np.float32(cutoff_ratio * n_freq)
        ks = [k for k in range(n_freq) if np.float32(k) <= lim]
        return ks[-1] if ks else -1
    etot = float(np.float32(spec.astype(np.float32).sum(dtype=np.float32)))
    cum = 0.0
    for i in range(n_freq):
        cum += float(spec[i])
        if cum >= cutoff_ratio * etot:
            return i
    return 0


def apply_fresca_to_score(score: np.ndarray, low_scale: float = 1.0, high_scale: float = 1.0,
                          cutoff_ratio: float = 0.5, cutoff_strategy: str = "energy",
                          timestep: Optional[float] = None, num_steps: Optional[int] = None) -> np.ndarray:
    """fresca.py:213-268 + frequency_scale :111-183 for score[B, L, C]."""
    h = high_scale
    if timestep is not None and num_steps is not None and high_scale > 1.0:
        tn = timestep / num_steps if num_steps > 0 else 0.0
        h = (1.0 - tn) * (high_scale - 1.0) + 1.0
    if low_scale == 1.0 and h == 1.0:
        return score
    L = score.shape[1]
    xf = np.fft.rfft(score.astype(np.float64), axis=1, norm="ortho")
    spec = np.abs(xf).astype(np.float32).mean(axis=(0, 2))
    rc = fresca_cutoff(spec, cutoff_ratio, cutoff_strategy)
    low = (np.arange(xf.shape[1]) <= rc).astype(np.float64)[None, :, None]
    xf = (low_scale * low + h * (1.0 - low)) * xf
    return np.fft.irfft(xf, n=L, axis=1, norm="ortho").astype(score.dtype)


# --------------------------------------------------------------------------- #
# Counter-based noise (the product's throughput-mode generator) restated
# --------------------------------------------------------------------------- #

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox-4x32-10 (Salmon et al. 2011).  counter [N,4] uint32, key [N,2] uint32."""
    c = counter.astype(np.uint32).copy()
    k = key.astype(np.uint32).copy()
    for _ in range(10):
        p0 = c[:, 0].astype(np.uint64) * _PHILOX_M0
        p1 = c[:, 2].astype(np.uint64) * _PHILOX_M1
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], axis=1)
        k = np.stack([k[:, 0] + _PHILOX_W0, k[:, 1] + _PHILOX_W1], axis=1)
    return c


def philox_normal(seed: int, stream: int, n: int) -> np.ndarray:
    """N(0,1) stream used by the CUDA sampler: element quad q=i//4 draws
    Philox(counter=(q, 0, stream_lo, stream_hi), key=(seed_lo, seed_hi)) and maps
    the four words to two Box-Muller pairs."""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    ctr = np.stack([q.astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                    np.full(nq, stream & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(nq, (stream >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=1)
    key = np.stack([np.full(nq, seed & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(nq, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=1)
    r = philox4x32_10(ctr, key).astype(np.float64)
    u = (r + 0.5) * (1.0 / 4294967296.0)  # (0,1)
    rad0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    rad1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    out = np.stack([rad0 * np.cos(2 * np.pi * u[:, 1]), rad0 * np.sin(2 * np.pi * u[:, 1]),
                    rad1 * np.cos(2 * np.pi * u[:, 3]), rad1 * np.sin(2 * np.pi * u[:, 3])], axis=1)
    return out.reshape(-1)[:n].astype(np.float32)


def event_intensity(crf, prev_crf, tau_0, step=1, dtype=np.float32):
    """E2CRFCache.compute_event_intensity (caching.py:524-559): prev None -> 1.0 at step 0 else 0.1; otherwise
    min(1, mean_{layer, token} ||crf - prev||_2 / tau_0)."""
    if prev_crf is None:
        return 0.1 if step > 0 else 1.0
    delta = np.abs(np.asarray(crf, dtype) - np.asarray(prev_crf, dtype))   # :541
    energy = np.sqrt((delta * delta).sum(-1, dtype=dtype))                # :544 torch.norm(dim=-1)
    avg = float(energy.mean(dtype=dtype))                                 # :547
    return min(1.0, avg / tau_0)                                          # :550


def spectral_density(x, apply_dft=True):
    """fourier.py:97-131 on ``x[B, L, C]``: |X_k|^2 per frequency bin from the real-packed DFT."""
    L = x.shape[1]
    xt = dft(x) if apply_dft else x                       # :109
    nr = n_real(L)                                        # :112
    re, im = xt[:, :nr, :], xt[:, nr:, :]                 # :113-114
    zero = np.zeros((x.shape[0], 1, x.shape[2]), xt.dtype)
    im = np.concatenate([zero, im], axis=1)               # :117-118
    if L % 2 == 0:
        im = np.concatenate([im, zero], axis=1)           # :121-122
    return re ** 2 + im ** 2                              # :129
