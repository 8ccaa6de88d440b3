"""Reading a reference (PyTorch-Lightning) checkpoint without Lightning, omegaconf or the reference installed.

``torch.load(..., weights_only=False)`` on such a file unpickles ``fdiff.schedulers.sde.VPScheduler`` (the
``noise_scheduler`` hyper-parameter, score_models.py:72) and Lightning containers (``AttributeDict``, callback
state).  The unpickler below maps the scheduler classes to this package's (same attribute names:
sde.py:22-23,58-64,105-106,184-185) and every other class of an absent package to an inert container, so only
``state_dict`` and ``hyper_parameters`` need to be meaningful.
"""

from __future__ import annotations

import pickle
from typing import Any, Dict

import torch

_ABSENT_OK = ("pytorch_lightning", "lightning", "lightning_fabric", "omegaconf", "fdiff", "torchmetrics", "wandb")


class _Inert(dict):
    """Stands in for an instance of a class whose package is not installed."""

    def __init__(self, *a: Any, **k: Any) -> None:
        try:
            super().__init__(*a, **k)
        except (TypeError, ValueError):
            super().__init__()

    def __setstate__(self, state: Any) -> None:
        if isinstance(state, dict):
            self.update(state)

    def __reduce_ex__(self, protocol):  # never re-pickled
        raise pickle.PicklingError("inert checkpoint container")


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        if module == "fdiff.schedulers.sde":
            from . import sde

            if hasattr(sde, name):
                return getattr(sde, name)
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            if module.split(".")[0] in _ABSENT_OK:
                return type(name, (_Inert,), {})
            raise


class _PickleModule:
    """The ``pickle_module`` protocol of ``torch.load``."""

    __name__ = "ffd_checkpoint_pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def load_lightning_checkpoint(path) -> Dict[str, Any]:
    ckpt = torch.load(str(path), map_location="cpu", weights_only=False, pickle_module=_PickleModule)
    if not isinstance(ckpt, dict) or "state_dict" not in ckpt:
        raise ValueError(f"{path}: not a Lightning checkpoint (no 'state_dict')")
    return ckpt
