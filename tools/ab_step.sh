# A/B of whole-step time for library variants on ONE box: tools/ab_step.sh build/variants/*.so  (run on the GPU box)
# Variants are selected through FFD_LIB; the in-tree library is never overwritten.
for rep in 1 2; do
for v in "$@"; do
  echo "== $v: $(FFD_LIB="$(readlink -f "$v")" python tools/time_step.py 2>&1 | tail -1)"
done
done
