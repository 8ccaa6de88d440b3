#!/usr/bin/env python
"""Mean time of a pure-cache step of the index schedule (ECG, B = 16384), CUDA events around one ffd_sample_steps
call:   python tools/time_step.py [--steps 100] [--reps 3]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import fastfourierdiffusion_b200 as F

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
model = bench.build_model("fp16").cuda().eval()
s = F.DiffusionSampler(model, a.batch, use_cache=True, noise="philox", cache_kwargs=dict(K=5, R=10))
model.noise_scheduler.set_timesteps(1000)
model.cache.reset()
X = s.run_steps(s.sample_prior(a.batch), 0, step_begin=0, step_end=3)
out = []
for r in range(a.reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    X = s.run_steps(X, 0, step_begin=3, step_end=3 + a.steps)
    e1.record()
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / a.steps)
print("ms per pure-cache step:", " ".join(f"{v:.3f}" for v in out))
