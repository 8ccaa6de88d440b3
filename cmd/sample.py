#!/usr/bin/env python
"""Sample from a score model -- entry point mirroring the reference's ``cmd/sample.py``.

    python cmd/sample.py num_samples=1000 num_diffusion_steps=1000 use_cache=true \\
        cache_kwargs="{'K': 5, 'R': 10}" sampler.sample_batch_size=500 [model_path=... model_id=...]

Same keys as cmd/conf/sample.yaml.  The flow is the reference's (cmd/sample.py:99-125): sample in the model's
domain, de-standardise, ``idft`` when the model works in the frequency domain, save ``samples.pt`` and
``results.yaml``.  Without a trained checkpoint (no datasets/checkpoints ship with this repo) a random-init
model of the named dataset's shape is used and the standardisation constants are the identity.  The quality
metrics of the reference (sliced Wasserstein, POT) are out of scope.
"""

from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import yaml

import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200.cli import SAMPLE_DEFAULTS, build_model, parse_overrides


def main(argv) -> None:
    cfg = parse_overrides(argv, SAMPLE_DEFAULTS)
    torch.manual_seed(cfg["random_seed"])
    model, ckpt = build_model(cfg)
    kwargs = dict(sample_batch_size=cfg["sampler.sample_batch_size"], noise=cfg["noise"], seed=cfg["random_seed"])
    if cfg["use_cache"]:
        kwargs.update(use_cache=True, cache_kwargs=dict(cfg["cache_kwargs"] or {}))
    sampler = F.DiffusionSampler(score_model=model, **kwargs)
    t0 = time.time()
    X = sampler.sample(num_samples=cfg["num_samples"], num_diffusion_steps=cfg["num_diffusion_steps"])
    torch.cuda.synchronize()
    dt = time.time() - t0
    if cfg["fourier_transform"]:
        X = F.idft(X)  # cmd/sample.py:111-113 (identity standardisation for random-init models)
    out_dir = cfg["output_dir"] or os.path.join(cfg["model_path"], str(cfg["model_id"] or "random_init"),
                                                "samples_cache" if cfg["use_cache"] else "samples")
    os.makedirs(out_dir, exist_ok=True)
    torch.save(X, os.path.join(out_dir, "samples.pt"))
    results = {"num_samples": int(X.shape[0]), "num_diffusion_steps": cfg["num_diffusion_steps"],
               "sampling_seconds": dt, "samples_per_second": X.shape[0] / dt, "checkpoint": ckpt,
               "use_cache": bool(cfg["use_cache"])}
    if cfg["use_cache"]:
        results["cache_stats"] = {k: float(v) for k, v in model.cache.get_cache_stats().items()}
    with open(os.path.join(out_dir, "results.yaml"), "w") as f:
        yaml.dump(results, f)
    print(yaml.dump(results))


if __name__ == "__main__":
    main(sys.argv[1:])
