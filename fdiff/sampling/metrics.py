"""``fdiff.sampling.metrics`` (reference: src/fdiff/sampling/metrics.py:28-216)."""

from fastfourierdiffusion_b200.metrics import MarginalWasserstein, MetricCollection, SlicedWasserstein

__all__ = ["MetricCollection", "SlicedWasserstein", "MarginalWasserstein"]
