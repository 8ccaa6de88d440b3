#!/usr/bin/env python
"""Per-step time of the bf16 sampling path (E2-CRF pure-cache steps and the full-recompute step) at the dataset
shapes BASELINE.json names: ECG, NASA-charge, NASDAQ, USDroughts (default model, random init).

    python tools/time_shapes.py [--rows 3000000]     # batch = rows / L per shape
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import fastfourierdiffusion_b200 as F

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=3_000_000)
a = ap.parse_args()
SHAPES = {"ECG": (187, 1), "NASA-charge": (251, 4), "NASDAQ": (252, 5), "USDroughts": (365, 13)}
out = {}
for name, (L, C) in SHAPES.items():
    torch.manual_seed(42)
    sched = F.VPScheduler(fourier_noise_scaling=True)
    model = F.ScoreModule(noise_scheduler=sched, fourier_noise_scaling=True, precision="fp16", max_len=L, n_channels=C,
                          d_model=72, num_layers=10, n_head=12).cuda().eval()
    sched.set_noise_scaling(L)
    B = max(128, a.rows // L // 128 * 128)
    s = F.DiffusionSampler(model, B, use_cache=True, noise="philox")
    sched.set_timesteps(1000)
    X = s.sample_prior(B)
    model.cache.reset()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    X = s.run_steps(X, 0, step_begin=0, step_end=1)  # warm-up + populate
    torch.cuda.synchronize()
    model.cache.reset()
    e[0].record()
    X = s.run_steps(X, 0, step_begin=0, step_end=1)
    e[1].record()
    X = s.run_steps(X, 0, step_begin=1, step_end=21)
    e[2].record()
    torch.cuda.synchronize()
    full, pure = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]) / 20
    out[name] = dict(L=L, C=C, batch=B, ms_full_step=round(full, 2), ms_pure_cache_step=round(pure, 2),
                     samples_per_s_1000_steps=round(B / (pure * 1000 / 1000 + (full - pure) / 1000) , 1))
    print(name, out[name], flush=True)
    del model, s, X
    torch.cuda.empty_cache()
print(json.dumps(out))
