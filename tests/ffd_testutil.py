"""Helpers shared by the test modules (imported as a top-level module: the
name ``tests`` is taken by another package on this image's sys.path)."""

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel(a, b):
    """max |a-b| / max |b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


# ---- product-side helpers (GPU tests) ---------------------------------------------
TINY = dict(max_len=16, n_channels=2, d_model=12, num_layers=2, n_head=2)
ECG = dict(max_len=187, n_channels=1, d_model=72, num_layers=10, n_head=12)


def tiny_state_dict():
    g = load_golden("tiny_vp_freq_nocache.npz")
    return {k[3:]: g[k] for k in g.files if k.startswith("w::")}


def make_scheduler(kind="vp", fns=True, **kw):
    import fastfourierdiffusion_b200 as F

    cls = F.VPScheduler if kind == "vp" else F.VEScheduler
    return cls(fourier_noise_scaling=fns, **kw)


def build_tiny_model(kind="vp", fns=True, precision="fp32", **skw):
    """Product ScoreModule carrying the tiny reference weights stored in the fixture."""
    import torch

    import fastfourierdiffusion_b200 as F

    sched = make_scheduler(kind, fns, **skw)
    m = F.ScoreModule(noise_scheduler=sched, fourier_noise_scaling=fns, precision=precision, **TINY)
    sd = {k: torch.from_numpy(np.array(v)) for k, v in tiny_state_dict().items()}
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    sched.set_noise_scaling(TINY["max_len"])
    return m.to("cuda").eval()


def build_seeded_model(dims, seed, kind="vp", fns=True, precision="fp32", distinct_layers=True, **skw):
    """Product ScoreModule random-initialised exactly like the reference under
    torch.manual_seed(seed) (+ oracle.ref_harness.redraw_layers, as make_golden.py does)."""
    import torch

    import fastfourierdiffusion_b200 as F
    from oracle import ref_harness as rh

    torch.manual_seed(seed)
    sched = make_scheduler(kind, fns, **skw)
    m = F.ScoreModule(noise_scheduler=sched, fourier_noise_scaling=fns, precision=precision, **dims)
    if distinct_layers:
        rh.redraw_layers(m)
    sched.set_noise_scaling(dims["max_len"])
    return m


def state_dict_checksums(model):
    sd = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in model.state_dict().items()}
    keys = sorted(sd)
    return keys, np.array([[np.abs(sd[k]).sum(), (sd[k] ** 2).sum()] for k in keys])
