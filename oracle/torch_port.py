"""Torch-CPU port of the reference sampling step, used ONLY as the timed CPU baseline.

TEST/BENCH INFRASTRUCTURE (see oracle/fdiff_oracle.py header): imported by
``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) and by the CPU tests
that pin it to the numpy oracle.  It restates the same algorithm as
``fdiff_oracle.py`` with the ATen CPU ops the reference itself dispatches to
(``F.linear``, ``torch.matmul``, ``F.softmax``, ``F.layer_norm``: SURVEY 2a), so
its speed on the host cores is representative of the reference's CPU sampler.
The reference itself cannot be imported on the GPU box (no upstream checkout
there), hence ``cpu_baseline.kind == "port"``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


class TorchPort:
    def __init__(self, sd: Dict[str, torch.Tensor], max_len: int, n_channels: int, n_head: int, sde_kind: str = "vp",
                 fourier_noise_scaling: bool = True, beta_min: float = 0.1, beta_max: float = 20.0,
                 sigma_min: float = 0.01, sigma_max: float = 50.0, device: str = "cpu"):
        # device="cuda" runs the same ATen op sequence as stock torch eager on the GPU (bench.py reports it next to
        # the CPU figure as "what the unmodified reference's ops would do on this GPU"); never used by the product
        self.dev = torch.device(device)
        self.sd = {k: v.detach().float().to(self.dev) for k, v in sd.items()}
        self.L, self.C, self.H = max_len, n_channels, n_head
        self.d = self.sd["embedder.weight"].shape[0]
        self.NL = sum(1 for k in self.sd if k.endswith("linear1.weight"))
        self.kind = sde_kind
        self.beta_min, self.beta_max, self.sigma_min, self.sigma_max = beta_min, beta_max, sigma_min, sigma_max
        G = torch.ones(max_len)
        if fourier_noise_scaling:  # sde.py:49-58
            G = 1 / math.sqrt(2) * G
            G[0] *= math.sqrt(2)
            if max_len % 2 == 0:
                G[max_len // 2] *= math.sqrt(2)
        self.G = G.to(self.dev)
        pos = self.sd["pos_encoder.embedding.weight"].clone()
        for _ in range(3):  # transformer.py:13-15 max-norm renorm
            torch.embedding_renorm_(pos, torch.arange(max_len, device=self.dev), math.sqrt(self.d), 2.0)
        self.pos = pos
        self.k_cache: Optional[torch.Tensor] = None
        self.v_cache: Optional[torch.Tensor] = None

    # -- model ---------------------------------------------------------------------
    def _layer(self, i: int, x: torch.Tensor, S: Optional[Sequence[int]]) -> torch.Tensor:
        p = f"backbone.layers.{i}."
        sd, d, H, L = self.sd, self.d, self.H, self.L
        B = x.shape[0]
        hd = d // H
        in_w, in_b = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]

        def heads(t):
            return t.view(B, -1, H, hd).transpose(1, 2)

        def tail(x, attn):
            a = F.linear(attn.transpose(1, 2).reshape(B, L, d), sd[p + "self_attn.out_proj.weight"],
                         sd[p + "self_attn.out_proj.bias"])
            x = F.layer_norm(x + a, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
            f = F.linear(torch.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                         sd[p + "linear2.weight"], sd[p + "linear2.bias"])
            return F.layer_norm(x + f, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)

        def attend(q, k, v):
            s = torch.matmul(q, k.transpose(-2, -1)) / (hd ** 0.5)
            return torch.matmul(F.softmax(s, dim=-1), v)

        full = S is None or len(S) == L or len(S) > 0.8 * L
        if full:  # cached_transformer.py:123-220
            qkv = F.linear(x, in_w, in_b)
            q, k, v = (heads(t) for t in qkv.split(d, dim=-1))
            y = tail(x, attend(q, k, v))
            if S is not None and len(S) == L:
                self._store(i, list(range(L)), heads(F.linear(y, in_w[d:2 * d], in_b[d:2 * d])),
                            heads(F.linear(y, in_w[2 * d:], in_b[2 * d:])))
            return y
        q = heads(F.linear(x, in_w[:d], in_b[:d]))
        if len(S) == 0:  # :237-258
            if self.k_cache is None:
                k = v = torch.zeros(B, H, L, hd, device=self.dev)
            else:
                k = self.k_cache[i].unsqueeze(0).expand(B, -1, -1, -1)
                v = self.v_cache[i].unsqueeze(0).expand(B, -1, -1, -1)
        else:  # :259-305
            S = sorted(S)
            k = torch.zeros(B, H, L, hd, device=self.dev)
            v = torch.zeros(B, H, L, hd, device=self.dev)
            rest = [j for j in range(L) if j not in set(S)]
            if rest and self.k_cache is not None:
                k[:, :, rest] = self.k_cache[i][:, rest].unsqueeze(0)
                v[:, :, rest] = self.v_cache[i][:, rest].unsqueeze(0)
            xs = x[:, S]
            kr, vr = heads(F.linear(xs, in_w[d:2 * d], in_b[d:2 * d])), heads(F.linear(xs, in_w[2 * d:], in_b[2 * d:]))
            k[:, :, S] = kr
            v[:, :, S] = vr
            self._store(i, S, kr, vr)
        return tail(x, attend(q, k, v))

    def _store(self, i: int, idx: List[int], k: torch.Tensor, v: torch.Tensor) -> None:
        if self.k_cache is None:  # caching.py:364-377
            self.k_cache = torch.zeros(self.NL, self.H, self.L, self.d // self.H, device=self.dev)
            self.v_cache = torch.zeros_like(self.k_cache)
        self.k_cache[i][:, idx] = k[0]
        self.v_cache[i][:, idx] = v[0]

    def score(self, X: torch.Tensor, t: float, S: Optional[Sequence[int]]) -> torch.Tensor:
        sd = self.sd
        h = F.linear(X, sd["embedder.weight"], sd["embedder.bias"]) + self.pos
        proj = torch.tensor(t, dtype=torch.float32, device=self.dev) * sd["time_encoder.W"] * 2 * math.pi
        e = torch.cat([torch.sin(proj), torch.cos(proj)])[: self.d]
        h = h + F.linear(e, sd["time_encoder.dense.weight"], sd["time_encoder.dense.bias"])
        for i in range(self.NL):
            h = self._layer(i, h, S)
        return F.linear(h, sd["unembedder.weight"], sd["unembedder.bias"])

    # -- SDE step (sde.py:143-163, 228-244), element-wise form ------------------------------
    def step(self, X: torch.Tensor, t: float, z: torch.Tensor, dt: torch.Tensor,
             S: Optional[Sequence[int]]) -> torch.Tensor:
        s = self.score(X, t, S)
        if self.kind == "vp":
            beta = self.beta_min + t * (self.beta_max - self.beta_min)
            diff = (math.sqrt(beta) * self.G)[None, :, None]
            drift = -0.5 * beta * X - diff * diff * s
        else:
            sdv = self.sigma_min * math.sqrt(2 * math.log(self.sigma_max / self.sigma_min)) * (
                self.sigma_max / self.sigma_min) ** t
            diff = (sdv * self.G)[None, :, None]
            drift = -(diff * diff * s)
        return X - drift * dt + torch.sqrt(dt) * (diff * z)


def recompute_set(step: int, K: int, R: int, L: int) -> List[int]:
    """caching.py:131-181."""
    if step == 0:
        return list(range(L))
    if step % (500 if R < 100 else R) == 0:
        return list(range(min(2 * min(K, L), L)))
    return []


def time_cached_steps(port: TorchPort, batch: int, steps: Sequence[int], n_total: int, K: int = 5, R: int = 10,
                      seed: int = 0):
    """Run the given schedule steps (absolute indices) on a fresh batch; returns seconds per step index."""
    import time

    g = torch.Generator().manual_seed(seed)
    ts = torch.linspace(1.0, 1e-5, n_total)
    dt = (ts[0] - ts[1]).to(port.dev)
    cuda = port.dev.type == "cuda"
    X = port.G[None, :, None] * torch.randn(batch, port.L, port.C, generator=g).to(port.dev)
    out = {}
    with torch.no_grad():
        for i in steps:
            z = torch.randn(batch, port.L, port.C, generator=g).to(port.dev)
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            X = port.step(X, float(ts[i]), z, dt, recompute_set(i, K, R, port.L))
            if cuda:
                torch.cuda.synchronize()
            out[i] = time.perf_counter() - t0
    return out, X
