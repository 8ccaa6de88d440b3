"""``fdiff.sampling.sampler`` (reference: src/fdiff/sampling/sampler.py:14-228)."""

from fastfourierdiffusion_b200.sampler import DiffusionSampler

__all__ = ["DiffusionSampler"]
