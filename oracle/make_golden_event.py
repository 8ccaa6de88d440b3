"""Reference vectors for E2CRFCache.compute_event_intensity (caching.py:524-559), from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/make_golden_event.py
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
    from fdiff.utils.caching import E2CRFCache  # type: ignore

    torch.manual_seed(7)
    d = {}
    for tag, (NL, L, dm, tau, scale) in {"tiny": (2, 16, 12, 0.1, 0.01), "ecg": (10, 187, 72, 0.1, 0.004),
                                          "sat": (2, 37, 60, 0.05, 1.0), "nasa": (2, 251, 72, 0.5, 0.03)}.items():
        c = E2CRFCache(num_layers=NL, max_len=L, device=torch.device("cpu"), tau_0=tau)
        prev = torch.randn(NL, L, dm)
        cur = prev + scale * torch.randn(NL, L, dm)
        d[f"{tag}_none0"] = np.float64(c.compute_event_intensity(cur, 0))
        d[f"{tag}_none5"] = np.float64(c.compute_event_intensity(cur, 5))
        c.current_step = 0
        c.update_crf(prev)
        d[f"{tag}_prev"], d[f"{tag}_cur"], d[f"{tag}_tau"] = prev.numpy(), cur.numpy(), np.float64(tau)
        d[f"{tag}_val"] = np.float64(c.compute_event_intensity(cur, 3))
    np.savez_compressed(os.path.join(OUT, "event_intensity.npz"), **d)
    print({k: float(v) for k, v in d.items() if k.endswith("_val") or "none" in k})


if __name__ == "__main__":
    main()
