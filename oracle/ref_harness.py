"""Import the UNMODIFIED reference classes from the upstream checkout.

TEST INFRASTRUCTURE ONLY (see oracle/fdiff_oracle.py header).  Works only where
the upstream checkout exists (this container: /root/reference); the GPU box has
no such path, so nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may
call :func:`load_reference` unless :func:`reference_available` is true.

``pytorch_lightning`` and ``diffusers`` are absent here; the reference imports
them at module scope (src/fdiff/models/score_models.py:3,7,9) but uses them only
for training, so two stub modules let the sampling path import unchanged.
"""

from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import Callable, Dict, List, Optional

REFERENCE_ROOT = os.environ.get("FFD_REFERENCE_ROOT", "/root/reference")


def reference_available(root: Optional[str] = None) -> bool:
    return os.path.isdir(os.path.join(root or REFERENCE_ROOT, "src", "fdiff"))


def _install_stubs() -> None:
    import torch.nn as nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

        class _Dummy:
            def __init__(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.LightningDataModule = _Dummy
        pl.Callback = _Dummy
        pl.Trainer = _Dummy
        plu = types.ModuleType("pytorch_lightning.utilities")
        plut = types.ModuleType("pytorch_lightning.utilities.types")
        plut.OptimizerLRScheduler = object
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = plu
        sys.modules["pytorch_lightning.utilities.types"] = plut
    if "diffusers" not in sys.modules:
        d = types.ModuleType("diffusers")
        do = types.ModuleType("diffusers.optimization")
        do.get_cosine_schedule_with_warmup = lambda **k: None
        sys.modules["diffusers"] = d
        sys.modules["diffusers.optimization"] = do


def load_reference(root: Optional[str] = None):
    """Return a namespace with the reference's hot-path classes (``root``: a directory holding ``src/fdiff``, default
    the upstream checkout; ``oracle/_ref`` is the copy ``oracle/ref_runner.vendor`` makes for the GPU box)."""
    root = root or REFERENCE_ROOT
    if not reference_available(root):
        raise RuntimeError(f"reference checkout not found at {root}")
    _install_stubs()
    src = os.path.join(root, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    # The upstream ``fdiff`` is a namespace package (no top-level __init__.py); this repository's drop-in ``fdiff``
    # package (a regular package) would win the import whatever the path order.  Bind the name to the upstream
    # directory explicitly, dropping any already-imported drop-in modules.
    for name in [m for m in sys.modules if m == "fdiff" or m.startswith("fdiff.")]:
        del sys.modules[name]
    pkg = types.ModuleType("fdiff")
    pkg.__path__ = [os.path.join(src, "fdiff")]
    sys.modules["fdiff"] = pkg
    from fdiff.models.score_models import ScoreModule  # type: ignore
    from fdiff.sampling.sampler import DiffusionSampler  # type: ignore
    from fdiff.schedulers.sde import VEScheduler, VPScheduler  # type: ignore
    from fdiff.utils.caching import E2CRFCache  # type: ignore
    from fdiff.utils.dataclasses import DiffusableBatch  # type: ignore
    from fdiff.utils.fourier import dft, idft  # type: ignore
    from fdiff.models.cached_transformer import CachedTransformerEncoderLayer  # type: ignore

    return types.SimpleNamespace(
        ScoreModule=ScoreModule, DiffusionSampler=DiffusionSampler, VPScheduler=VPScheduler,
        VEScheduler=VEScheduler, E2CRFCache=E2CRFCache, DiffusableBatch=DiffusableBatch,
        dft=dft, idft=idft, CachedTransformerEncoderLayer=CachedTransformerEncoderLayer)


def build_reference_model(ref, *, seed: int, max_len: int, n_channels: int, d_model: int = 72,
                          num_layers: int = 10, n_head: int = 12, sde: str = "vp",
                          fourier_noise_scaling: bool = True, sde_kwargs: Optional[dict] = None,
                          distinct_layers: bool = True):
    """Random-init reference model under ``torch.manual_seed(seed)``.

    ``nn.TransformerEncoder`` deep-copies one layer, so a fresh model has NL
    identical layers; ``distinct_layers`` re-draws layers 1.. (same init
    functions, continuing the same RNG stream) so layer indexing mistakes cannot
    hide.  The product's ScoreModule mirror performs the same steps.
    """
    import torch

    torch.manual_seed(seed)
    kw = dict(sde_kwargs or {})
    sched = (ref.VPScheduler if sde == "vp" else ref.VEScheduler)(
        fourier_noise_scaling=fourier_noise_scaling, **kw)
    model = ref.ScoreModule(n_channels=n_channels, max_len=max_len, noise_scheduler=sched,
                            fourier_noise_scaling=fourier_noise_scaling, d_model=d_model,
                            num_layers=num_layers, n_head=n_head)
    if distinct_layers:
        redraw_layers(model)
    sched.set_noise_scaling(max_len)
    model.eval()
    return model


def redraw_layers(model) -> None:
    """Give layers 1.. their own weights and make LayerNorm affine params and
    biases non-trivial (default init leaves them at 1/0, which would hide bugs)."""
    import torch

    with torch.no_grad():
        for i, layer in enumerate(model.backbone.layers):
            for name, p in layer.named_parameters():
                if i > 0 and p.dim() == 2:
                    torch.nn.init.xavier_uniform_(p)
                elif "norm" in name and name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn_like(p))
                elif p.dim() == 1:
                    p.copy_(0.05 * torch.randn_like(p))


@contextlib.contextmanager
def record_randn(log: List):
    """Log every ``torch.randn`` / ``torch.randn_like`` draw (sde.py:85,157,238)."""
    import torch

    o_randn, o_like = torch.randn, torch.randn_like

    def randn(*a, **k):
        z = o_randn(*a, **k)
        log.append(z.detach().clone())
        return z

    def randn_like(x, **k):
        z = o_like(x, **k)
        log.append(z.detach().clone())
        return z

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = o_randn, o_like


@contextlib.contextmanager
def record_steps(sampler, log: List[Dict]):
    """Log (X_t, t, recompute set, X_{t-1}) around reverse_diffusion_step
    (sampler.py:48-103)."""
    orig: Callable = sampler.reverse_diffusion_step

    def wrapped(batch, step=0, recompute_tokens=None):
        out = orig(batch, step=step, recompute_tokens=recompute_tokens)
        log.append(dict(X=batch.X.detach().clone(), t=float(batch.timesteps[0]),
                        S=None if recompute_tokens is None else sorted(recompute_tokens),
                        Xn=out.detach().clone()))
        return out

    sampler.reverse_diffusion_step = wrapped
    try:
        yield
    finally:
        sampler.reverse_diffusion_step = orig


def state_dict_numpy(model) -> Dict[str, "object"]:
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()
            if not k.startswith("cached_backbone")}
