"""``fdiff.utils.wasserstein`` (reference: src/fdiff/utils/wasserstein.py:12-199)."""

from fastfourierdiffusion_b200.metrics import WassersteinDistances

__all__ = ["WassersteinDistances"]
