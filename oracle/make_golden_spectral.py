"""Reference vectors for fdiff.utils.fourier.spectral_density (fourier.py:97-131), from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/make_golden_spectral.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    rh.load_reference()
    from fdiff.utils.fourier import dft, spectral_density  # type: ignore

    torch.manual_seed(11)
    d = {}
    for L, C, B in [(24, 3, 5), (187, 1, 4), (101, 2, 3), (252, 5, 2)]:
        x = torch.randn(B, L, C)
        d[f"x_{L}"] = x.numpy()
        d[f"sd_{L}"] = spectral_density(x).numpy()
        d[f"sdf_{L}"] = spectral_density(dft(x), apply_dft=False).numpy()
    np.savez_compressed(os.path.join(OUT, "spectral.npz"), **d)
    print({k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
