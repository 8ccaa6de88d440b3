"""Cycle timeline of CTA 0 around its local tile 2 of the fused layer kernel.
Build:  FFD_NVCC_FLAGS=-DFFD_LAYER_STAMPS FFD_LIB_OUT=build/variants/stamps.so python fastfourierdiffusion_b200/_build.py
Run:    FFD_LIB=build/variants/stamps.so python tools/layer_stamps.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
import fastfourierdiffusion_b200 as F
from fastfourierdiffusion_b200 import _lib
lib = _lib.load()
model = bench.build_model("fp16").cuda().eval()
B = 16384
L, d, H = model.max_len, model.d_model, model.n_head
M = B * L
NL = 10
arena, packed = model.device_weights()
kc = torch.randn(NL, H, L, d // H, device="cuda"); vc = torch.randn(NL, H, L, d // H, device="cuda")
ws = model.workspace(B)
x = torch.randn(M, d, device="cuda"); y = torch.empty_like(x)
st = _lib.stream_handle()
def layer():
    _lib.check(lib.ffd_layer_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc), _lib.ptr(vc),
                                       _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel(), M, st))
layer(); torch.cuda.synchronize()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.ffd_debug_set_attn_stamps.argtypes = [ctypes.c_void_p]
lib.ffd_debug_set_attn_stamps(buf.data_ptr())
layer(); torch.cuda.synchronize()
lib.ffd_debug_set_attn_stamps(None)
t = buf.cpu().numpy()
t0 = int(t[0]) if t[0] else int(t[t > 0].min())
f = lambda v: int(v - t0) if v else None
NV = 24
print("softmax (warp 16 = even heads, warp 24 = odd heads) per virtual head of tile 2: s_full seen, p_full arrived")
for n in range(NV): print(n, [f(v) for v in t[2 * n:2 * n + 2]])
print("attention MMA warp per virtual head: p_full seen, PV + next S issued")
for n in range(NV): print(n, [f(v) for v in t[500 + 2 * n:502 + 2 * n]])
print("tile warp 8, period 3: start | F1(2) done | Q(4) done | x_full arrived | X(5) staged | [y_full seen] | F2(2) done")
print([f(t[i]) for i in (200, 201, 202, 203, 204, 205, 206)], " F1: obf_full seen, residual landed:", [f(t[207]), f(t[208])])
print("FFN MMA warp tile 2: x_full seen", f(t[299]), " GEMM2(j) issued:")
print([f(v) for v in t[300:332]])
print("FFN epilogue warp 12, tile 2, per half-slice: h_full seen, r_full arrived")
print([(f(t[400 + 2 * j]), f(t[401 + 2 * j])) for j in range(32)])
print("FFN MMA warp, tile 2, j = 8, 9: loop top, waits done, fence done, step issued")
print([f(v) for v in t[600:604]], [f(v) for v in t[610:614]])
print("attention MMA warp, virtual heads 10, 11: loop top, [p_full seen], fence done, step issued")
print([f(t[642]), f(t[520]), f(t[640]), f(t[641])], [f(t[646]), f(t[522]), f(t[644]), f(t[645])])
print("softmax quad 1 p_full arrivals per virtual head: group 0, group 1")
for n in range(NV): print(n, [f(v) for v in t[700 + 2 * n:702 + 2 * n]])
print("read-out warp (quad 0) per virtual head: o_full seen, o_empty arrived")
for n in range(NV): print(n, [f(v) for v in t[760 + 2 * n:762 + 2 * n]])
print("attention MMA warp, virtual heads 10 / 11, serial waits: top, p_full, o_empty, v_full, k_full")
print([f(v) for v in t[900:905]], [f(v) for v in t[908:913]])
