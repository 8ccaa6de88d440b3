"""``fdiff.utils.fresca`` (reference: src/fdiff/utils/fresca.py:111-268)."""

from fastfourierdiffusion_b200.fresca import apply_fresca_to_score, frequency_scale

__all__ = ["frequency_scale", "apply_fresca_to_score"]
