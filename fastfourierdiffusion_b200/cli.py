"""Minimal ``key=value`` configuration for the entry points in ``cmd/`` (Hydra is not required).

The key names and defaults are the reference's (cmd/conf/sample.yaml:1-12, cmd/conf/sampler/default.yaml,
cmd/conf/score_model/default.yaml:3-5, cmd/conf/score_model/noise_scheduler/{vpsde,vesde}.yaml).
"""

from __future__ import annotations

import ast
import os
import re
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


def find_best_checkpoint(ckpt_dir: str) -> Optional[str]:
    """Best checkpoint of a reference training run: lowest val_loss among ``*epoch=N-val_loss=X.ckpt``
    (utils/extraction.py:79-98)."""
    if not os.path.isdir(ckpt_dir):
        return None
    best, best_loss = None, float("inf")
    for f in sorted(os.listdir(ckpt_dir)):
        m = re.match(r"(.*?)epoch=(\d+)-val_loss=(\d+\.\d+)\.ckpt$", f)
        if m is not None and float(m.group(3)) < best_loss:
            best, best_loss = os.path.join(ckpt_dir, f), float(m.group(3))
    return best


# ---- the run directory of a trained model (cmd/sample.py:31-75) ---------------------------------------------
def resolve_config(cfg: Any, root: Optional[dict] = None) -> Any:
    """``train_config.yaml`` is written by ``OmegaConf.save`` with its interpolations unresolved
    (cmd/train.py:43): resolve ``${a.b}`` references against the root; resolver calls such as
    ``${hydra:runtime.cwd}`` become the current directory."""
    root = cfg if root is None else root

    def lookup(path: str) -> Any:
        node: Any = root
        for part in path.split("."):
            node = node[part]
        return resolve_config(node, root)

    if isinstance(cfg, dict):
        return {k: resolve_config(v, root) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [resolve_config(v, root) for v in cfg]
    if isinstance(cfg, str) and "${" in cfg:
        whole = re.fullmatch(r"\$\{([^${}:]+)\}", cfg)
        if whole:
            return lookup(whole.group(1))
        cfg = re.sub(r"\$\{hydra:runtime\.cwd\}", os.getcwd().replace("\\", "/"), cfg)
        return re.sub(r"\$\{([^${}:]+)\}", lambda m: str(lookup(m.group(1))), cfg)
    return cfg


def instantiate(node: Any, **extra: Any) -> Any:
    """The subset of ``hydra.utils.instantiate`` the sampling entry points use: ``_target_`` (a dotted path, the
    reference's ``fdiff.*`` names resolve through the ``fdiff`` package of this repository), ``_partial_``,
    nested nodes."""
    import functools
    import importlib

    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return {k: instantiate(v) for k, v in node.items()}
    mod, _, name = node["_target_"].rpartition(".")
    fn = getattr(importlib.import_module(mod), name)
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
    kwargs.update(extra)
    return functools.partial(fn, **kwargs) if node.get("_partial_") else fn(**kwargs)


def load_run(model_path: str, model_id: str) -> Dict[str, Any]:
    """What cmd/sample.py:31-75 reads from ``<model_path>/<model_id>``: ``train_config.yaml`` (datamodule, domain,
    standardisation) and the best checkpoint.  Missing pieces raise ``FileNotFoundError`` as the reference does --
    there is no silent fall-back to a random-init model when a model id is given."""
    import yaml

    run_dir = os.path.join(model_path, str(model_id))
    if not os.path.isdir(run_dir):
        avail = sorted(d for d in os.listdir(model_path) if os.path.isdir(os.path.join(model_path, d))) \
            if os.path.isdir(model_path) else []
        raise FileNotFoundError(f"Model directory not found: {run_dir}\nAvailable model IDs: {', '.join(avail[:10])}")
    cfg_path = os.path.join(run_dir, "train_config.yaml")
    if not os.path.exists(cfg_path):
        raise FileNotFoundError(f"Training config not found: {cfg_path}\nThis model may not have been trained yet.")
    with open(cfg_path) as f:
        train_cfg = resolve_config(yaml.safe_load(f))
    ckpt = find_best_checkpoint(os.path.join(run_dir, "checkpoints"))
    if ckpt is None:
        raise FileNotFoundError(f"no '*epoch=N-val_loss=X.ckpt' checkpoint under {run_dir}/checkpoints")
    return {"run_dir": run_dir, "train_cfg": train_cfg, "checkpoint": ckpt}


def build_model(cfg: Dict[str, Any]):
    """ScoreModule on the GPU.  With ``model_id``: the trained run -- architecture, SDE and domain from the
    checkpoint's hyper-parameters and ``train_config.yaml`` (cmd/sample.py:50-75), never from the CLI defaults.
    Without: random init under ``random_seed`` on the named dataset's shape (same init as the reference under the
    same seed; no datasets / checkpoints ship with this repository).  Returns (model, checkpoint path or None)."""
    from . import ScoreModule

    if cfg.get("model_id"):
        run = load_run(cfg["model_path"], cfg["model_id"])
        model = ScoreModule.load_from_checkpoint(run["checkpoint"], precision=cfg.get("precision"))
        cfg["_run"] = run
        return model.to("cuda").eval(), run["checkpoint"]
    max_len, n_channels = DATASET_SHAPES[cfg["datamodule"]]
    torch.manual_seed(cfg["random_seed"])
    sched = build_scheduler(cfg["noise_scheduler"], bool(cfg["fourier_transform"]))
    model = ScoreModule(n_channels=n_channels, max_len=max_len, noise_scheduler=sched,
                        fourier_noise_scaling=bool(cfg["fourier_transform"]), d_model=cfg["score_model.d_model"],
                        num_layers=cfg["score_model.num_layers"], n_head=cfg["score_model.n_head"],
                        precision=cfg.get("precision"))
    sched.set_noise_scaling(max_len)
    return model.to("cuda").eval(), None
