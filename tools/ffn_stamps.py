"""Cycle timeline of one tile pair of the FFN kernel (build with FFD_NVCC_FLAGS=-DFFD_FFN_STAMPS)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes, torch, bench
from fastfourierdiffusion_b200 import _lib
lib = _lib.load()
model = bench.build_model("fp16").cuda().eval()
M = 16384 * 187
arena, packed = model.device_weights()
x = torch.randn(M, 72, device="cuda"); y = torch.empty_like(x); st = _lib.stream_handle()
def ffn(): _lib.check(lib.ffd_ffn_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(x), _lib.ptr(y), None, M, st))
ffn(); torch.cuda.synchronize()
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
lib.ffd_debug_set_attn_stamps.argtypes = [ctypes.c_void_p]
lib.ffd_debug_set_attn_stamps(buf.data_ptr())
ffn(); torch.cuda.synchronize()
lib.ffd_debug_set_attn_stamps(None)
t = buf.cpu().numpy(); t0 = t[t > 0].min()
f = lambda v: int(v - t0) if v else None
print("slice: MMA warp(tile 0): p_full seen, GEMM2 issued, GEMM1(next) issued | epilogue warp 12: h_full seen, p_full arrive")
for j in range(16): print(j, [f(v) for v in t[j*8:j*8+5]])
print("tile warp 4 (pair 2): start, X(3,0) staged, LN(2,0) done", [f(v) for v in t[200:203]])
