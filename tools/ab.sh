# A/B timing of library variants on ONE box: tools/ab.sh build/variants/*.so   (run on the GPU box)
# Build variants locally, e.g.:  FFD_NVCC_FLAGS=-DX python fastfourierdiffusion_b200/_build.py && cp fastfourierdiffusion_b200/libffd_b200.so build/variants/x.so
cp fastfourierdiffusion_b200/libffd_b200.so /tmp/ffd_keep.so
for v in "$@"; do
  cp "$v" fastfourierdiffusion_b200/libffd_b200.so
  echo "== $v: $(python tools/time_kernels.py 2>&1 | grep -E 'attn|ffn' | tr '\n' ' ')"
done
cp /tmp/ffd_keep.so fastfourierdiffusion_b200/libffd_b200.so
