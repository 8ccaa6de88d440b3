"""``fdiff.utils.fourier`` (reference: src/fdiff/utils/fourier.py:8-131)."""

from fastfourierdiffusion_b200.fourier import dft, idft, spectral_density

__all__ = ["dft", "idft", "spectral_density"]
