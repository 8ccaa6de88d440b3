# A/B timing of library variants on ONE box: tools/ab.sh build/variants/*.so   (run on the GPU box)
# Build variants locally, e.g.:  FFD_NVCC_FLAGS=-DX python fastfourierdiffusion_b200/_build.py && cp fastfourierdiffusion_b200/libffd_b200.so build/variants/x.so
# The variant is selected through FFD_LIB (read by fastfourierdiffusion_b200/_lib.py): the in-tree library is never
# overwritten, so an interrupted run cannot leave a variant in its place.
for v in "$@"; do
  echo "== $v: $(FFD_LIB="$(readlink -f "$v")" python tools/time_kernels.py 2>&1 | grep -E 'attn|ffn|layer' | tr '\n' ' ')"
done
