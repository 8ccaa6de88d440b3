#!/usr/bin/env python
"""Launch one of the dominant kernels a few times on ECG-sized inputs (for ncu captures).

    python tools/prof_kernels.py --kernel ffn|attn|layer [--batch 16384] [--reps 3]

ffn / attn launch the block's C-ABI entry point directly (attn = the pure-cache attention block on a shared K/V cache);
layer runs whole sampler steps (step 0 = stock attention in chunks, then pure-cache steps).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import bench
from fastfourierdiffusion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="ffn")
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

import __graft_entry__ as e

e.build()
lib = _lib.load()
model = bench.build_model("fp16").cuda().eval()
M = a.batch * 187
st = _lib.stream_handle()
if a.kernel == "ffn":
    arena, packed = model.device_weights()
    x = torch.randn(M, 72, device="cuda")
    y = torch.empty_like(x)
    for _ in range(a.reps):
        _lib.check(lib.ffd_ffn_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(x), _lib.ptr(y), None, M, st))
elif a.kernel == "attn":
    arena, packed = model.device_weights()
    L, d, H = model.max_len, model.d_model, model.n_head
    x = torch.randn(M, d, device="cuda")
    y = torch.empty_like(x)
    kc = torch.randn(10, H, L, d // H, device="cuda")
    vc = torch.randn(10, H, L, d // H, device="cuda")
    ws = model.workspace(a.batch)
    for _ in range(a.reps):
        _lib.check(lib.ffd_attn_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc), _lib.ptr(vc),
                                          _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel(), M, st))
else:
    import fastfourierdiffusion_b200 as F

    s = F.DiffusionSampler(model, a.batch, use_cache=True, noise="philox")
    model.noise_scheduler.set_timesteps(1000)
    X = s.sample_prior(a.batch)
    model.cache.reset()
    s.run_steps(X, 0, step_begin=0, step_end=1 + a.reps)
torch.cuda.synchronize()
print("ok")
