"""``fdiff.utils.dataclasses`` (reference: src/fdiff/utils/dataclasses.py:7-18)."""

from fastfourierdiffusion_b200.batch import DiffusableBatch

__all__ = ["DiffusableBatch"]
