#!/usr/bin/env python
"""Sample from a score model -- entry point mirroring the reference's ``cmd/sample.py``.

    python cmd/sample.py num_samples=1000 num_diffusion_steps=1000 use_cache=true \\
        cache_kwargs="{'K': 5, 'R': 10}" sampler.sample_batch_size=500 [model_path=... model_id=...]

Same keys as cmd/conf/sample.yaml.  The flow is the reference's (cmd/sample.py:99-125): sample in the model's
domain, de-standardise, ``idft`` when the model works in the frequency domain, save ``samples.pt`` and
``results.yaml``.  With ``model_id`` the run directory must hold ``train_config.yaml`` and a checkpoint
(``FileNotFoundError`` otherwise, as in the reference): architecture, SDE, domain and standardisation come from
them.  Without ``model_id`` (no datasets/checkpoints ship with this repo) a random-init model of the named
dataset's shape is used and the standardisation constants are the identity.  The quality
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
from fastfourierdiffusion_b200.cli import SAMPLE_DEFAULTS, build_model, instantiate, parse_overrides


def main(argv) -> None:
    cfg = parse_overrides(argv, SAMPLE_DEFAULTS)
    torch.manual_seed(cfg["random_seed"])
    model, ckpt = build_model(cfg)
    fourier_transform, feature_std, feature_mean = bool(cfg["fourier_transform"]), None, None
    if ckpt is not None:
        # cmd/sample.py:66-70,107-109: domain and standardisation belong to the TRAINING run, not to this call
        datamodule = instantiate(cfg["_run"]["train_cfg"]["datamodule"])
        datamodule.prepare_data()
        datamodule.setup()
        fourier_transform = bool(datamodule.fourier_transform)
        if datamodule.standardize:
            feature_mean, feature_std = datamodule.feature_mean_and_std
    kwargs = dict(sample_batch_size=cfg["sampler.sample_batch_size"], noise=cfg["noise"], seed=cfg["random_seed"])
    if cfg["use_cache"]:
        kwargs.update(use_cache=True, cache_kwargs=dict(cfg["cache_kwargs"] or {}))
    sampler = F.DiffusionSampler(score_model=model, **kwargs)
    t0 = time.time()
    # de-standardise + idft on the device, per batch, before the copy to the host (cmd/sample.py:107-113)
    X = sampler.sample_time_domain(num_samples=cfg["num_samples"], num_diffusion_steps=cfg["num_diffusion_steps"],
                                   feature_std=feature_std, feature_mean=feature_mean,
                                   fourier_transform=fourier_transform)
    torch.cuda.synchronize()
    dt = time.time() - t0
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
