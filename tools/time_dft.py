#!/usr/bin/env python
"""Time of the boundary transforms (dft / idft incl. the fused de-standardise) on device-resident batches."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fastfourierdiffusion_b200 as F

out = {}
for name, (B, L, C) in {"ECG": (16384, 187, 1), "NASA-charge": (11904, 251, 4), "USDroughts": (8192, 365, 13)}.items():
    x = torch.randn(B, L, C, device="cuda")
    std, mean = torch.rand(L, C, device="cuda") + 0.5, torch.randn(L, C, device="cuda")
    for fn, call in (("dft", lambda: F.dft(x)), ("idft_destandardise", lambda: F.idft(x, std, mean))):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f"{name}_{fn}"] = {"B": B, "L": L, "C": C, "ms": ms, "gflops": 2.0 * L * L * C * B / ms / 1e6}
print(json.dumps(out))
