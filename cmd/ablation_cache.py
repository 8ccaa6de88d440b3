#!/usr/bin/env python
"""E2-CRF cache ablation -- entry point mirroring the reference's ``cmd/ablation_cache.py``.

    python cmd/ablation_cache.py num_samples=20 num_diffusion_steps=5 [fourier_transform=false] \
        [sampler.sample_batch_size=1]

Runs the reference's six configurations (ablation_cache.py:126-214): baseline without the cache, the default
E2-CRF cache, ``tau_warn=0`` ("no event trigger"), ``R=tau_warn=999999`` ("no error feedback": with R >= 100 the
refresh interval is R itself, caching.py:164-181, so only step 0 recomputes), ``tau_0=0`` ("no energy weighting")
and the "naive" setting ``K=5, tau_0=0, tau_warn=0, R=999999``; prints the reference's summary table and writes
``ablation_results.json`` with its keys (elapsed_time, speedup, cache_stats).  The reference samples at batch size
1; that default is kept, any batch size can be given.  ``fourier_transform=false`` selects the time-domain model
(BASELINE.json config 5: no frequency noise scaling, no DFT).
"""

from __future__ import annotations

import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200.cli import SAMPLE_DEFAULTS, build_model, parse_overrides

CONFIGS = [  # (name, use_cache, cache_kwargs): ablation_cache.py:126-214
    ("Baseline", False, None),
    ("E2-CRF (full)", True, {}),
    ("No event trigger", True, {"tau_warn": 0.0}),
    ("No error feedback", True, {"R": 999999, "tau_warn": 999999}),
    ("No energy weighting", True, {"tau_0": 0.0}),
    ("Naive caching", True, {"K": 5, "tau_0": 0.0, "tau_warn": 0.0, "R": 999999}),
]


def run_ablation(model, num_samples, num_diffusion_steps, batch_size, config_name, cache_kwargs, use_cache, noise,
                 seed):
    """ablation_cache.py:24-75."""
    sampler = F.DiffusionSampler(score_model=model, sample_batch_size=batch_size, use_cache=use_cache,
                                 cache_kwargs=cache_kwargs, noise=noise, seed=seed)
    if not use_cache:
        model.disable_caching()
    torch.manual_seed(seed)
    torch.cuda.synchronize()
    t0 = time.time()
    samples = sampler.sample(num_samples=num_samples, num_diffusion_steps=num_diffusion_steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    stats = model.cache.get_cache_stats() if (use_cache and model.cache is not None) else {}
    return {"config_name": config_name, "elapsed_time": dt, "metrics": {}, "cache_stats": stats,
            "finite": bool(torch.isfinite(samples).all()), "n": int(samples.shape[0])}


def main(argv) -> None:
    defaults = dict(SAMPLE_DEFAULTS)
    defaults.update({"num_samples": 20, "num_diffusion_steps": 5, "sampler.sample_batch_size": 1})
    cfg = parse_overrides(argv, defaults)
    model, _ = build_model(cfg)
    n, steps, bs = cfg["num_samples"], cfg["num_diffusion_steps"], cfg["sampler.sample_batch_size"]
    run_ablation(model, min(n, bs), min(steps, 3), bs, "warm-up", None, False, cfg["noise"], cfg["random_seed"])
    print("=" * 80 + "\nE2-CRF Caching Ablation Study\n" + "=" * 80)
    results = []
    for name, use_cache, kw in CONFIGS:
        r = run_ablation(model, n, steps, bs, name, kw, use_cache, cfg["noise"], cfg["random_seed"])
        results.append(r)
        print(f"{name}: {r['elapsed_time']:.3f}s  speedup {results[0]['elapsed_time'] / r['elapsed_time']:.2f}x")
    base = results[0]["elapsed_time"]
    print("\n" + "=" * 80 + "\nAblation Study Results\n" + "=" * 80)
    print(f"{'Configuration':<30} {'Time (s)':<12} {'Speedup':<10} {'Cache Hit Ratio':<15}")
    print("-" * 80)
    for r in results:
        hit = r["cache_stats"].get("cache_hit_ratio", 0.0) if r["cache_stats"] else 0.0
        print(f"{r['config_name']:<30} {r['elapsed_time']:<12.3f} {base / r['elapsed_time']:<10.2f} {hit:<15.2%}")
    out = cfg["output_dir"] or "ablation_results"
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "ablation_results.json"), "w") as f:
        json.dump({r["config_name"]: {"elapsed_time": r["elapsed_time"], "speedup": base / r["elapsed_time"],
                                      "cache_stats": r["cache_stats"], "finite": r["finite"]} for r in results},
                  f, indent=2)
    print(f"\nResults saved to {os.path.join(out, 'ablation_results.json')}")


if __name__ == "__main__":
    main(sys.argv[1:])
