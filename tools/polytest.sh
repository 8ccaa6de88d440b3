for m in 0x00 0x88 0x92 0xAA; do
  FFD_NVCC_FLAGS="-DFFD_POLY_PAIRS=${m}u" python fastfourierdiffusion_b200/_build.py >/dev/null 2>&1
  python - <<PY
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch, bench
from fastfourierdiffusion_b200 import _lib
lib=_lib.load()
model=bench.build_model("bf16").cuda().eval()
import fastfourierdiffusion_b200 as F
B=16384; M=B*187
s=F.DiffusionSampler(model,B,use_cache=True,noise="philox"); model.noise_scheduler.set_timesteps(1000)
X=s.sample_prior(B); model.cache.reset(); s.run_steps(X,0,step_begin=0,step_end=1)
arena,packed=model.device_weights(); kc,vc=model.cache._k_store,model.cache._v_store; ws=model.workspace(B)
x=torch.randn(M,72,device="cuda"); y=torch.empty_like(x); st=_lib.stream_handle()
def attn(): _lib.check(lib.ffd_attn_cached_ln(model.dims,1,_lib.ptr(arena),_lib.ptr(packed),3,_lib.ptr(kc),_lib.ptr(vc),_lib.ptr(x),_lib.ptr(y),_lib.ptr(ws),ws.numel(),M,st))
for _ in range(3): attn()
torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): attn()
b.record(); torch.cuda.synchronize()
print("mask ${m}: attn ms", a.elapsed_time(b)/10)
PY
done
