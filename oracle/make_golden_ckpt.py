"""Freeze a tiny run directory in the REFERENCE's on-disk format (tests/golden/ref_run/).

TEST INFRASTRUCTURE ONLY.  Runs where the upstream checkout exists (this container).  A Lightning checkpoint is
a pickled dict with ``state_dict`` and ``hyper_parameters`` (score_models.py:72 ``save_hyperparameters``); the
hyper-parameters hold the UNMODIFIED reference scheduler object (pickled as ``fdiff.schedulers.sde.VEScheduler``)
inside Lightning's ``AttributeDict`` container (pickled as ``lightning_fabric.utilities.data.AttributeDict`` --
Lightning itself is not installed here, so a class of that module path and name stands in for it: the bytes a
loader sees name the same class).  ``train_config.yaml`` is written the way ``OmegaConf.save`` leaves it
(cmd/train.py:43): interpolations unresolved.

    python oracle/make_golden_ckpt.py
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_run")
MODEL_ID = "tinyve01"
DIMS = dict(n_channels=1, max_len=16, d_model=12, num_layers=2, n_head=2)

TRAIN_CONFIG = """random_seed: 42
fourier_transform: true
standardize: true
score_model:
  _target_: fdiff.models.score_models.ScoreModule
  _partial_: true
  d_model: 12
  num_layers: 2
  n_head: 2
  lr_max: 0.001
  fourier_noise_scaling: ${fourier_transform}
  likelihood_weighting: false
  noise_scheduler:
    _target_: fdiff.schedulers.sde.VEScheduler
    eps: 1.0e-05
    sigma_min: 0.01
    sigma_max: 2
    fourier_noise_scaling: ${score_model.fourier_noise_scaling}
datamodule:
  _target_: fdiff.dataloaders.datamodules.SyntheticDatamodule
  data_dir: ${hydra:runtime.cwd}/data
  random_seed: ${random_seed}
  fourier_transform: ${fourier_transform}
  standardize: ${standardize}
  batch_size: 64
  max_len: 16
  num_samples: 64
"""


def main() -> None:
    ref = rh.load_reference()
    # Lightning's container for hyper-parameters, by module path and name
    mod = types.ModuleType("lightning_fabric.utilities.data")

    class AttributeDict(dict):
        pass

    AttributeDict.__module__ = "lightning_fabric.utilities.data"
    AttributeDict.__qualname__ = "AttributeDict"
    mod.AttributeDict = AttributeDict
    for name in ("lightning_fabric", "lightning_fabric.utilities"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lightning_fabric.utilities.data"] = mod

    torch.manual_seed(7)
    sched = ref.VEScheduler(sigma_min=0.01, sigma_max=2.0, fourier_noise_scaling=True)
    model = ref.ScoreModule(noise_scheduler=sched, fourier_noise_scaling=True, **DIMS)
    rh.redraw_layers(model)
    model.eval()
    hp = AttributeDict(noise_scheduler=sched, fourier_noise_scaling=True, num_training_steps=1000, lr_max=1e-3,
                       likelihood_weighting=False, **DIMS)
    ckpt = {"epoch": 1, "global_step": 10, "pytorch-lightning_version": "2.1.0",
            "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
            "hparams_name": "kwargs", "hyper_parameters": hp}
    run = os.path.join(OUT, MODEL_ID)
    os.makedirs(os.path.join(run, "checkpoints"), exist_ok=True)
    torch.save(ckpt, os.path.join(run, "checkpoints", "epoch=1-val_loss=0.50.ckpt"))
    torch.save({"state_dict": {}}, os.path.join(run, "checkpoints", "epoch=0-val_loss=0.90.ckpt"))
    with open(os.path.join(run, "train_config.yaml"), "w") as f:
        f.write(TRAIN_CONFIG)
    # what the reference computes from this checkpoint: scores of one batch, and 6 reverse steps with given noise
    g = torch.Generator().manual_seed(11)
    X = torch.randn(3, DIMS["max_len"], DIMS["n_channels"], generator=g)
    t = torch.full((3,), 0.37)
    with torch.no_grad():
        score = model(ref.DiffusableBatch(X=X, y=None, timesteps=t))
    np.savez(os.path.join(run, "expected.npz"), X=X.numpy(), t=t.numpy(), score=score.numpy(),
             sigma_max=np.float32(sched.sigma_max))
    print("wrote", run)


if __name__ == "__main__":
    main()
