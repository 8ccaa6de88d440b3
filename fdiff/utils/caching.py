"""``fdiff.utils.caching`` (reference: src/fdiff/utils/caching.py:19-636)."""

from fastfourierdiffusion_b200.caching import E2CRFCache

__all__ = ["E2CRFCache"]
