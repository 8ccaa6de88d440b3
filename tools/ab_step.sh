# A/B of whole-step time for library variants on ONE box: tools/ab_step.sh build/variants/*.so  (run on the GPU box)
cp fastfourierdiffusion_b200/libffd_b200.so /tmp/ffd_keep.so
for rep in 1 2; do
for v in "$@"; do
  cp "$v" fastfourierdiffusion_b200/libffd_b200.so
  echo "== $v: $(python tools/time_step.py 2>&1 | tail -1)"
done
done
cp /tmp/ffd_keep.so fastfourierdiffusion_b200/libffd_b200.so
