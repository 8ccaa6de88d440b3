#!/usr/bin/env python
"""Time of the three step kinds of the index schedule at ECG shapes: full recompute (step 0), pure cache, and the
mixed refresh step (step 500: first 2K tokens recomputed).   python tools/time_mixed.py [--batch 16384]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import fastfourierdiffusion_b200 as F

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
a = ap.parse_args()
model = bench.build_model("fp16").cuda().eval()
B = a.batch
s = F.DiffusionSampler(model, B, use_cache=True, noise="philox", cache_kwargs=dict(K=5, R=10))
model.noise_scheduler.set_timesteps(1000)
model.cache.reset()
X = s.sample_prior(B)
X = s.run_steps(X, 0, step_begin=0, step_end=2)
out = {}
for name, (b, e) in (("pure_cache_step", (2, 12)), ("mixed_step_500", (500, 501)), ("mixed_step_500_again", (500, 501)), ("full_step_0", (0, 1))):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    X = s.run_steps(X, 0, step_begin=b, step_end=e)
    e1.record()
    torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / (e - b)
print(json.dumps(out))
