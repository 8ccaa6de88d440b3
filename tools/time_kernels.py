#!/usr/bin/env python
"""Time the two dominant kernels alone (CUDA events, 10 back-to-back launches) at ECG shapes.

    python tools/time_kernels.py [--batch 16384] [--len 187]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import bench
import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
a = ap.parse_args()
lib = _lib.load()
model = bench.build_model("fp16").cuda().eval()
B = a.batch
L = model.max_len
M = B * L
s = F.DiffusionSampler(model, B, use_cache=True, noise="philox")
model.noise_scheduler.set_timesteps(1000)
X = s.sample_prior(B)
model.cache.reset()
s.run_steps(X, 0, step_begin=0, step_end=1)
arena, packed = model.device_weights()
kc, vc = model.cache._k_store, model.cache._v_store
ws = model.workspace(B)
x = torch.randn(M, model.dims.d_model, device="cuda")
y = torch.empty_like(x)
st = _lib.stream_handle()


def attn():
    _lib.check(lib.ffd_attn_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc), _lib.ptr(vc),
                                      _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel() * ws.element_size(), M, st))


def ffn():
    _lib.check(lib.ffd_ffn_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(x), _lib.ptr(y), None, M, st))


def layer():
    _lib.check(lib.ffd_layer_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc), _lib.ptr(vc),
                                       _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel() * ws.element_size(), M, st))


for name, fn in (("attn", attn), ("ffn", ffn), ("layer", layer)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.4f} ms per launch (M = {M})")
