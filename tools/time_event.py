#!/usr/bin/env python
"""Step time of the device-decided E2-CRF schedule (event_driven=True) against the reference's index schedule,
ECG shapes, CUDA events around one ffd_sample_steps call each.

    python tools/time_event.py [--batch 16384] [--steps 60] [--out profiles/x.json]
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
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--out", default=None)
a = ap.parse_args()
model = bench.build_model("fp16").cuda().eval()
B, n = a.batch, a.steps
rows = []
cases = [("index schedule (reference): K=5 R=10", dict(K=5, R=10))]
for tau0 in (1e30, 3.0, 0.3, 0.03, 0.0):
    cases.append((f"event-driven tau_0={tau0:g} K=5 R=50 tau_warn=0.5 probe=0.02",
                  dict(K=5, R=50, tau_0=tau0, tau_warn=0.5, event_driven=True, event_probe_ratio=0.02)))
for name, kw in cases:
    s = F.DiffusionSampler(model, B, use_cache=True, noise="philox", cache_kwargs=kw)
    model.noise_scheduler.set_timesteps(1000)
    model.cache.reset()
    X = s.sample_prior(B)
    X = s.run_steps(X, 0, step_begin=0, step_end=3)  # step 0 (full recompute) + warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    X = s.run_steps(X, 0, step_begin=3, step_end=3 + n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    st = model.cache.get_event_stats() if kw.get("event_driven") else {}
    rows.append({"schedule": name, "ms_per_step": ms, "finite": bool(torch.isfinite(X).all()), **st})
    print(json.dumps(rows[-1]))
if a.out:
    json.dump({"batch": B, "steps_timed": n, "rows": rows}, open(a.out, "w"), indent=1)
