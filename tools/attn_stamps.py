"""Cycle timeline of one token tile of the attention kernel (build with FFD_NVCC_FLAGS=-DFFD_ATTN_STAMPS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes, torch, bench
import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200 import _lib
lib = _lib.load()
model = bench.build_model("fp16").cuda().eval()
B = 16384
s = F.DiffusionSampler(model, B, use_cache=True, noise="philox")
model.noise_scheduler.set_timesteps(1000)
X = s.sample_prior(B); model.cache.reset()
s.run_steps(X, 0, step_begin=0, step_end=2)
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
lib.ffd_debug_set_attn_stamps.argtypes = [ctypes.c_void_p]
lib.ffd_debug_set_attn_stamps(buf.data_ptr())
s.run_steps(X, 0, step_begin=1, step_end=2)
torch.cuda.synchronize()
lib.ffd_debug_set_attn_stamps(None)
t = buf.cpu().numpy()
t0 = t[t > 0].min()
f = lambda v: int(v - t0) if v else None
NV = 12
print("score warp per S issued: S issued, slots done (cycles from t0)")
for h in range(NV): print(h, [f(v) for v in t[h*2:h*2+2]])
print("P.V warp per head: p_full seen, PV issued")
for h in range(NV): print(h, [f(v) for v in t[64+h*2:64+h*2+2]])
print("softmax warp 12 per head: loop top, before next s_full wait, after it, p_full arrive")
for h in range(NV): print(h, [f(v) for v in t[128+h*4:128+h*4+4]])
print("read-out warp 4 per head: o_full seen, o_empty arrive")
for h in range(NV): print(h, [f(v) for v in t[256+h*2:256+h*2+2]])
print("tile warp 8 (phase 2): start, finish(1) done, prepare(3) done | prepare(3): x staged, t_full seen")
print([f(v) for v in t[400:405]])

if os.environ.get("FFD_ATTN_QUARTERS") == "1":
    NVH = 24
    for g in range(4):
        print(f"v12 softmax group {g} (warp {12+4*g}) per virtual head: top, s_full seen, max done, p_full arrive")
        for h in range(NVH): print(h, [f(v) for v in t[512+(g*32+h)*4:512+(g*32+h)*4+4]])
    for g in range(4):
        print(f"v12 MMA warp {g}: before wait, after wait, PV issued, S issued")
        for h in range(NVH): print(h, [f(v) for v in t[1024+(g*32+h)*4:1024+(g*32+h)*4+4]])
