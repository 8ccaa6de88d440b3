"""Minimal ``key=value`` configuration for the entry points in ``cmd/`` (Hydra is not required).

The key names and defaults are the reference's (cmd/conf/sample.yaml:1-12, cmd/conf/sampler/default.yaml,
cmd/conf/score_model/default.yaml:3-5, cmd/conf/score_model/noise_scheduler/{vpsde,vesde}.yaml).
"""

from __future__ import annotations

import ast
import os
from typing import Any, Dict, List, Optional

import torch

# (max_len, n_channels) of the reference's datamodules (datamodules.py:194-203,349-353,406-410,469-476,530-538)
DATASET_SHAPES = {"ecg": (187, 1), "nasdaq": (252, 5), "nasa_charge": (251, 4), "nasa_discharge": (134, 5),
                  "droughts": (365, 13), "mimiciii": (24, 40), "synthetic": (100, 1)}

SAMPLE_DEFAULTS: Dict[str, Any] = {
    "model_path": "lightning_logs", "model_id": None, "num_samples": 100, "num_diffusion_steps": 100,
    "random_seed": 42, "use_cache": False, "cache_kwargs": {}, "sampler.sample_batch_size": 50,
    # used when no trained checkpoint is given (random-init model on synthetic shapes)
    "datamodule": "ecg", "fourier_transform": True, "noise_scheduler": "vpsde",
    "score_model.d_model": 72, "score_model.num_layers": 10, "score_model.n_head": 12,
    "precision": None, "noise": "torch", "output_dir": None,
}


def parse_overrides(argv: List[str], defaults: Dict[str, Any]) -> Dict[str, Any]:
    cfg = dict(defaults)
    for a in argv:
        if "=" not in a:
            raise SystemExit(f"expected key=value, got {a!r}")
        k, v = a.split("=", 1)
        if k not in cfg:
            raise SystemExit(f"unknown key {k!r}; known keys: {sorted(cfg)}")
        try:
            val = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            val = {"true": True, "false": False, "null": None, "none": None}.get(v.lower(), v)
        cfg[k] = val
    return cfg


def build_scheduler(name: str, fourier_noise_scaling: bool):
    from . import VEScheduler, VPScheduler

    if name == "vpsde":  # vpsde.yaml: beta_min .1, beta_max 20
        return VPScheduler(beta_min=0.1, beta_max=20.0, fourier_noise_scaling=fourier_noise_scaling)
    if name == "vesde":  # vesde.yaml: sigma_min .01, sigma_max 2
        return VEScheduler(sigma_min=0.01, sigma_max=2.0, fourier_noise_scaling=fourier_noise_scaling)
    raise SystemExit(f"unknown noise_scheduler {name!r} (vpsde | vesde)")


def find_checkpoint(model_path: str, model_id: Optional[str]) -> Optional[str]:
    """Best checkpoint of a reference training run: lowest val_loss in ``epoch=..-val_loss=..ckpt``
    (utils/extraction.py:79-98)."""
    if not model_id:
        return None
    d = os.path.join(model_path, str(model_id), "checkpoints")
    if not os.path.isdir(d):
        return None
    best, best_loss = None, float("inf")
    for f in os.listdir(d):
        if f.endswith(".ckpt") and "val_loss=" in f:
            try:
                loss = float(f.split("val_loss=")[1].rsplit(".ckpt", 1)[0])
            except ValueError:
                continue
            if loss < best_loss:
                best, best_loss = os.path.join(d, f), loss
    return best


def build_model(cfg: Dict[str, Any]):
    """ScoreModule on the GPU: weights from a reference checkpoint's ``state_dict`` when one is found,
    else random init under ``random_seed`` (same init as the reference under the same seed)."""
    from . import ScoreModule

    max_len, n_channels = DATASET_SHAPES[cfg["datamodule"]]
    torch.manual_seed(cfg["random_seed"])
    sched = build_scheduler(cfg["noise_scheduler"], bool(cfg["fourier_transform"]))
    model = ScoreModule(n_channels=n_channels, max_len=max_len, noise_scheduler=sched,
                        fourier_noise_scaling=bool(cfg["fourier_transform"]), d_model=cfg["score_model.d_model"],
                        num_layers=cfg["score_model.num_layers"], n_head=cfg["score_model.n_head"],
                        precision=cfg.get("precision"))
    ckpt = find_checkpoint(cfg["model_path"], cfg["model_id"])
    if ckpt is not None:
        sd = torch.load(ckpt, map_location="cpu", weights_only=False)
        sd = sd.get("state_dict", sd)
        sd = {k: v for k, v in sd.items() if not k.startswith("cached_backbone")}
        model.load_state_dict(sd, strict=True)
    sched.set_noise_scaling(max_len)
    return model.to("cuda").eval(), ckpt
