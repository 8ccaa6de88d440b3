"""v12 (quarter streams, an experiment: build with FFD_NVCC_FLAGS=-DFFD_BUILD_ATTN_QUARTERS) against v11 of the attention
kernel: outputs and time.  FFD_LIB=build/variants/quarters.so python tools/ab_attn.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ffd_testutil import ECG, build_seeded_model, rel
from fastfourierdiffusion_b200 import _lib
lib = _lib.load()
lib.ffd_debug_set_attn_quarters.argtypes = [ctypes.c_int]
SH = {"ecg": dict(ECG), "nasa": dict(ECG, max_len=251, n_channels=4), "droughts": dict(ECG, max_len=365, n_channels=13)}
def run(model, x, kc, vc, B, mode, reps=1):
    lib.ffd_debug_set_attn_quarters(mode)
    arena, packed = model.device_weights(); ws = model.workspace(B); y = torch.empty_like(x)
    def call():
        _lib.check(lib.ffd_attn_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc), _lib.ptr(vc),
                                          _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel(), x.shape[0], _lib.stream_handle()))
    call(); torch.cuda.synchronize()
    ms = None
    if reps > 1:
        for _ in range(2): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): call()
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / reps
    lib.ffd_debug_set_attn_quarters(-1)
    return y, ms
for name, Bs in (("ecg", (1, 70, 1024)), ("nasa", (5, 300)), ("droughts", (3, 200))):
    model = build_seeded_model(SH[name], 42, precision="fp16").to("cuda").eval()
    d, L, H = model.d_model, model.max_len, model.n_head
    for B in Bs:
        torch.manual_seed(B)
        x = torch.randn(B * L, d, device="cuda") * 1.5
        kc = torch.randn(10, H, L, d // H, device="cuda"); vc = torch.randn(10, H, L, d // H, device="cuda")
        a, _ = run(model, x, kc, vc, B, 0); b, _ = run(model, x, kc, vc, B, 1)
        print(f"{name} B={B}: v12 vs v11 rel {rel(b.cpu().numpy(), a.cpu().numpy()):.2e} finite {bool(torch.isfinite(b).all())}", flush=True)
model = build_seeded_model(SH["ecg"], 42, precision="fp16").to("cuda").eval()
B = 16384; L = 187
x = torch.randn(B * L, 72, device="cuda"); kc = torch.randn(10, 12, L, 6, device="cuda"); vc = torch.randn(10, 12, L, 6, device="cuda")
for mode in (0, 1, 0, 1):
    _, ms = run(model, x, kc, vc, B, mode, reps=10)
    print(f"ECG B=16384 attention block {'v12' if mode else 'v11'}: {ms:.4f} ms (incl. K/V image pack)", flush=True)
