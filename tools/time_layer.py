"""Time the fused layer kernel alone: python tools/time_layer.py [batch]   (FFD_LIB selects a library variant)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from fastfourierdiffusion_b200 import _lib
lib = _lib.load()
model = bench.build_model("fp16").cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L, d, H = model.max_len, model.d_model, model.n_head
M = B * L
arena, packed = model.device_weights()
kc = torch.randn(10, H, L, d // H, device="cuda"); vc = torch.randn(10, H, L, d // H, device="cuda")
ws = model.workspace(B)
x = torch.randn(M, d, device="cuda"); y = torch.empty_like(x)
def layer():
    _lib.check(lib.ffd_layer_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc), _lib.ptr(vc),
                                       _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel(), M, _lib.stream_handle()))
for _ in range(2): layer()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): layer()
e1.record(); torch.cuda.synchronize()
print(f"layer: {e0.elapsed_time(e1) / 5:.4f} ms per launch (M = {M})")
