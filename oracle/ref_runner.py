"""The UNMODIFIED reference sampler as the timed CPU baseline (``cpu_baseline.kind == "reference"``).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/fdiff_oracle.py header).  The reference is pure Python: there is nothing to
compile, but it does not exist on the GPU box.  ``vendor()`` -- run by ``__graft_entry__.build()`` in the container that
has the upstream checkout -- copies the Python files of its sampling path, untouched, from ``/root/reference/src/fdiff``
into ``oracle/_ref/src/fdiff`` (git-ignored build output that travels with the snapshot like the built ``.so``; no
reference source enters the repository's history).  ``time_reference_steps`` then drives the reference's own classes
through its own per-step public API in the order its ``DiffusionSampler.sample`` does (sampler.py:157-210): timestep
tensor, ``DiffusableBatch``, ``cache.determine_recompute_set``, ``reverse_diffusion_step`` -- on the host cores, on the
1000-step time grid, for the schedule steps asked for.  ``pytorch_lightning`` / ``diffusers`` (training-only imports of
the reference) are stubbed by ``oracle/ref_harness.py``.
"""

from __future__ import annotations

import os
import shutil
import time
from typing import Dict, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
UPSTREAM = os.environ.get("FFD_REFERENCE_ROOT", "/root/reference")
# the modules the sampling path imports (src/fdiff/sampling/sampler.py and what it pulls in)
FILES = ["models/__init__.py", "models/score_models.py", "models/transformer.py", "models/cached_transformer.py",
         "sampling/__init__.py", "sampling/sampler.py", "schedulers/sde.py", "utils/__init__.py", "utils/caching.py",
         "utils/dataclasses.py", "utils/fourier.py", "utils/fresca.py", "utils/losses.py", "utils/caching_tensor.py"]


def vendor() -> Optional[str]:
    """Copy the reference's sampling-path modules into oracle/_ref (no-op where the upstream checkout is absent)."""
    src = os.path.join(UPSTREAM, "src", "fdiff")
    if not os.path.isdir(src):
        return REF_DIR if available() else None
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(REF_DIR, "src", "fdiff", rel)
        if os.path.exists(a):
            os.makedirs(os.path.dirname(b), exist_ok=True)
            shutil.copyfile(a, b)
    return REF_DIR


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "src", "fdiff", "sampling", "sampler.py"))


def time_reference_steps(dims: Dict[str, int], batch: int, steps: Sequence[int], n_total: int = 1000, seed: int = 42,
                         cache_kwargs: Optional[dict] = None, n_threads: Optional[int] = None
                         ) -> Tuple[Dict[int, float], "object"]:
    """Seconds per schedule step (absolute indices, in order) of the reference's cached sampling loop on the CPU."""
    import torch

    from oracle import ref_harness as rh

    if n_threads:
        torch.set_num_threads(n_threads)
    ref = rh.load_reference(REF_DIR)
    model = rh.build_reference_model(ref, seed=seed, sde="vp", fourier_noise_scaling=True, distinct_layers=False, **dims)
    sampler = ref.DiffusionSampler(score_model=model, sample_batch_size=batch, use_cache=True,
                                   cache_kwargs=dict(cache_kwargs or {}))
    model.eval()
    model.noise_scheduler.set_timesteps(n_total)
    out: Dict[int, float] = {}
    with torch.no_grad():
        X = sampler.sample_prior(batch)
        cache = model.cache
        cache.reset()
        ts = list(model.noise_scheduler.timesteps)
        for i in steps:
            t0 = time.perf_counter()
            timesteps = torch.full((batch,), ts[i].item(), dtype=torch.float32, device=model.device, requires_grad=False)
            b = ref.DiffusableBatch(X=X, y=None, timesteps=timesteps)
            cache.current_step = i
            S = cache.determine_recompute_set(x_tilde=X, event_intensity=0.1, step=i)
            X = sampler.reverse_diffusion_step(b, step=i, recompute_tokens=S)
            out[i] = time.perf_counter() - t0
    return out, X
