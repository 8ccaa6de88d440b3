#!/usr/bin/env python
"""Small-batch latency of the cached step loop, direct launches against CUDA-graph replay (FFD_GRAPH=0 / 1), wall
clock around DiffusionSampler.run_steps so that host launch overhead is included:
    python tools/time_graph.py            -> JSON, batches 1 and 64, 200 steps of the 1000-step grid (K=5, R=10)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
import fastfourierdiffusion_b200 as F

out = {}
model = bench.build_model("fp16").cuda().eval()
model.noise_scheduler.set_timesteps(1000)
for B in (1, 64):
    for mode in ("0", "1"):
        os.environ["FFD_GRAPH"] = mode
        s = F.DiffusionSampler(model, B, use_cache=True, noise="philox", cache_kwargs=dict(K=5, R=10))
        model.cache.reset()
        X = s.run_steps(s.sample_prior(B), 0, step_begin=0, step_end=21)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            X = s.run_steps(X, 0, step_begin=21, step_end=221)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / 200 * 1e3)
        out[f"B{B}_{'graph' if mode == '1' else 'direct'}_ms_per_step"] = round(best, 4)
    out[f"B{B}_samples_per_s_1000_steps_graph"] = round(B / out[f"B{B}_graph_ms_per_step"], 1)
print(json.dumps(out))
