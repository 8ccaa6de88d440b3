#!/usr/bin/env python
"""BASELINE.json configs[0]: ECG, batch 64, default model, VP-SDE, no cache (the reference's own CPU-runnable case):
ms per reverse-diffusion step on the GPU, stock branch and E2-CRF pure-cache branch.

    python tools/time_small_batch.py [--batch 64]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import bench
import fastfourierdiffusion_b200 as F

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
out = {"batch": a.batch}
for prec in ("fp16", "fp32"):
    model = bench.build_model(prec).cuda().eval()
    model.noise_scheduler.set_timesteps(1000)
    for use_cache in (False, True):
        s = F.DiffusionSampler(model, a.batch, use_cache=use_cache, noise="philox")
        if not use_cache:
            model.disable_caching()
        X = s.sample_prior(a.batch)
        if use_cache:
            model.cache.reset()
        X = s.run_steps(X, 0, step_begin=0, step_end=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        X = s.run_steps(X, 0, step_begin=3, step_end=103)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        out[f"{prec}_{'cache' if use_cache else 'nocache'}_ms_per_step"] = round(ms, 4)
        out[f"{prec}_{'cache' if use_cache else 'nocache'}_samples_per_s_1000_steps"] = round(a.batch / ms, 1)
print(json.dumps(out))
