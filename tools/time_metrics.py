#!/usr/bin/env python
"""Time the device Wasserstein metrics at the reference's default setting (metrics/default.yaml: 1000 random
directions) on ECG-sized sample sets, next to the numpy restatement on a bounded number of directions.

    python tools/time_metrics.py [--n 16384] [--m 14000] [--dirs 1000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import fastfourierdiffusion_b200 as F
from oracle import metrics_oracle as MO

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--m", type=int, default=14000)
ap.add_argument("--dirs", type=int, default=1000)
a = ap.parse_args()
rng = np.random.default_rng(0)
L, C = 187, 1
orig = torch.from_numpy(rng.normal(size=(a.m, L, C)).astype(np.float32)).cuda()
gen = torch.from_numpy((rng.normal(size=(a.n, L, C)) * 1.1 + 0.1).astype(np.float32)).cuda()
sw = F.SlicedWasserstein(orig, random_seed=42, num_directions=a.dirs)
mw = F.MarginalWasserstein(orig, random_seed=42)
sw(gen), mw(gen)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
r_s = sw(gen)
torch.cuda.synchronize()
t1 = time.perf_counter()
r_m = mw(gen)
torch.cuda.synchronize()
t2 = time.perf_counter()
# device-only share (CUDA events around the kernels of one sliced evaluation; directions pre-drawn)
wd = F.WassersteinDistances(orig, gen, seed=42)
dirs = np.stack(wd.get_random_directions(a.dirs))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
wd._distances(dirs)
e1.record()
torch.cuda.synchronize()
nd = 8
o, g = orig.cpu().numpy(), gen.cpu().numpy()
c0 = time.perf_counter()
_, d_ref = MO.sliced_wasserstein(o, g, 42, nd)
c1 = time.perf_counter()
out = {"n_generated": a.n, "n_original": a.m, "features": L * C, "directions": a.dirs,
       "sliced_wall_s": t1 - t0, "marginal_wall_s": t2 - t1, "sliced_device_ms": e0.elapsed_time(e1),
       "numpy_restatement_s_per_direction": (c1 - c0) / nd,
       "numpy_restatement_s_extrapolated": (c1 - c0) / nd * a.dirs,
       "sliced_wasserstein_mean": r_s["sliced_wasserstein_mean"], "marginal_wasserstein_mean": r_m["marginal_wasserstein_mean"]}
print(json.dumps(out))
