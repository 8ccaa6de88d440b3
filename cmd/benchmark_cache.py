#!/usr/bin/env python
"""E2-CRF cache benchmark -- entry point mirroring the reference's ``cmd/benchmark_cache.py``.

    python cmd/benchmark_cache.py num_samples=20 num_diffusion_steps=100 [sampler.sample_batch_size=1]

Times ``DiffusionSampler.sample`` (wall clock around the call, as benchmark_cache.py:92-99 does) without
the cache and with the cache for the reference's sweeps: default kwargs, K in {0,3,5,10}, R in {5,10,20,50},
tau_0 in {.05,.1,.2,.5}, freq_decomp_interval in {5,10,20,50} (benchmark_cache.py:172-391; only K and R change
the schedule), plus the FreSca sections (cache + FreSca, benchmark_cache.py:203-212, and the high_scale sweep,
:394-422).  The reference runs at batch size 1; that default is kept, any batch size can be given.
Writes ``benchmark_results.csv``.
"""

from __future__ import annotations

import csv
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200.cli import SAMPLE_DEFAULTS, build_model, parse_overrides


def benchmark_sampling(model, num_samples, num_steps, batch_size, use_cache, cache_kwargs, noise, seed,
                       use_fresca=False, fresca_kwargs=None):
    sampler = F.DiffusionSampler(score_model=model, sample_batch_size=batch_size, use_cache=use_cache,
                                 cache_kwargs=cache_kwargs, noise=noise, seed=seed, use_fresca=use_fresca,
                                 **(fresca_kwargs or {}))
    if not use_cache:
        model.disable_caching()
    torch.manual_seed(seed)
    torch.cuda.synchronize()
    t0 = time.time()
    X = sampler.sample(num_samples=num_samples, num_diffusion_steps=num_steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    stats = model.cache.get_cache_stats() if (use_cache and model.cache is not None) else {}
    return dt, X, stats


def main(argv) -> None:
    defaults = dict(SAMPLE_DEFAULTS)
    defaults.update({"num_samples": 20, "num_diffusion_steps": 100, "sampler.sample_batch_size": 1})
    cfg = parse_overrides(argv, defaults)
    model, _ = build_model(cfg)
    n, steps, bs = cfg["num_samples"], cfg["num_diffusion_steps"], cfg["sampler.sample_batch_size"]
    configs = [("baseline (no cache)", False, None, None), ("E2-CRF default", True, {}, None),
               ("E2-CRF + FreSca", True, {"use_fresca_in_cache": True},
                {"fresca_high_scale": 1.5, "fresca_cutoff_ratio": 0.5})]
    configs += [(f"K={k}", True, {"K": k}, None) for k in (0, 3, 5, 10)]
    configs += [(f"R={r}", True, {"R": r}, None) for r in (5, 10, 20, 50)]
    configs += [(f"tau_0={t}", True, {"tau_0": t}, None) for t in (0.05, 0.1, 0.2, 0.5)]
    configs += [(f"freq_decomp_interval={i}", True, {"freq_decomp_interval": i}, None) for i in (5, 10, 20, 50)]
    configs += [(f"FreSca h={h}", True, {"use_fresca_in_cache": True},
                 {"fresca_high_scale": h, "fresca_cutoff_ratio": 0.5}) for h in (1.0, 1.2, 1.5, 2.0)]
    benchmark_sampling(model, min(n, bs), 3, bs, False, None, cfg["noise"], cfg["random_seed"])  # warm-up
    rows, base = [], None
    for name, use_cache, kw, fk in configs:
        dt, X, stats = benchmark_sampling(model, n, steps, bs, use_cache, kw, cfg["noise"], cfg["random_seed"],
                                          use_fresca=fk is not None, fresca_kwargs=fk)
        base = dt if base is None else base
        rows.append({"config": name, "time_s": dt, "speedup": base / dt, "samples_per_s": X.shape[0] / dt,
                     "cache_hit_ratio": stats.get("cache_hit_ratio", 0.0), "cache_ratio": stats.get("cache_ratio", 0.0),
                     "finite": bool(torch.isfinite(X).all())})
        print(f"{name:28s} {dt:8.3f} s  x{base / dt:5.2f}  hit {rows[-1]['cache_hit_ratio'] * 100:5.1f}%")
    out = cfg["output_dir"] or "."
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "benchmark_results.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0]))
        w.writeheader()
        w.writerows(rows)


if __name__ == "__main__":
    main(sys.argv[1:])
