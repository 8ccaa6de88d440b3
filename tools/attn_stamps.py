import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes, torch, bench
import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200 import _lib
lib = _lib.load()
model = bench.build_model("bf16").cuda().eval()
B = 16384
s = F.DiffusionSampler(model, B, use_cache=True, noise="philox")
model.noise_scheduler.set_timesteps(1000)
X = s.sample_prior(B); model.cache.reset()
s.run_steps(X, 0, step_begin=0, step_end=2)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
lib.ffd_debug_set_attn_stamps.argtypes = [ctypes.c_void_p]
lib.ffd_debug_set_attn_stamps(buf.data_ptr())
s.run_steps(X, 0, step_begin=1, step_end=2)
torch.cuda.synchronize()
lib.ffd_debug_set_attn_stamps(None)
t = buf.cpu().numpy()
t0 = t[t > 0].min()
print("MMA thread per head: p_full, o_empty, v_full, PV issued, S issued (cycles from t0)")
for h in range(12): print(h, [int(v - t0) if v else None for v in t[h*8:h*8+5]])
print("softmax warp 2 per head: s_full, collected, pass1 done, barrier passed, p_full arrive")
for h in range(12): print(h, [int(v - t0) if v else None for v in t[128+h*8:128+h*8+5]])
print("collect_o (warp 2): after o_full wait, after ld")
for h in range(12): print(h, [int(v - t0) if v else None for v in t[256+h*4:256+h*4+3]])
print("tile phases (warp 2): loop start, t_full, prev out_full, x staged, q_full, qbf arrive | obf arrive, LN start (warp 10 = group 2), LN done, store read")
print([int(v - t0) if v else None for v in t[300:310]])
