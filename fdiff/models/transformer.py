"""``fdiff.models.transformer`` (reference: src/fdiff/models/transformer.py:8-29,61-91): parameter containers;
the lookups run inside the embed / time-embedding kernels."""

from fastfourierdiffusion_b200.score_model import GaussianFourierProjection, PositionalEncoding

__all__ = ["PositionalEncoding", "GaussianFourierProjection"]
