"""``fdiff.utils.extraction`` -- the three helpers ``cmd/sample.py`` uses (reference:
src/fdiff/utils/extraction.py:58-112)."""

from __future__ import annotations

from pathlib import Path
from typing import Any

from fastfourierdiffusion_b200.cli import find_best_checkpoint


def get_model_type(cfg: Any):
    """extraction.py:58-76: the class named by ``cfg.score_model._target_``."""
    from fdiff.models import score_models

    target = cfg["score_model"]["_target_"]
    prefix = "fdiff.models.score_models."
    if not target.startswith(prefix) or not hasattr(score_models, target[len(prefix):]):
        raise NotImplementedError(f"Model class {target} not implemented yet.")
    return getattr(score_models, target[len(prefix):])


def get_best_checkpoint(checkpoint_path: Path) -> Path:
    """extraction.py:79-98: lowest ``val_loss`` among ``*epoch=N-val_loss=X.ckpt``."""
    best = find_best_checkpoint(str(checkpoint_path))
    if best is None:
        raise FileNotFoundError(f"no '*epoch=N-val_loss=X.ckpt' checkpoint under {checkpoint_path}")
    return Path(best)


def dict_to_str(d: Any) -> str:
    """extraction.py:101-112."""
    return "".join(f"{k}: {v}\n" for k, v in dict(d).items())
