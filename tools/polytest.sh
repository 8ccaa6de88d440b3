# sweep the FMA-pipe share of the softmax exponentials (run on the GPU box)
for m in ${MASKS:-0x00 0x01 0x11 0x49 0x55}; do
  FFD_NVCC_FLAGS="-DFFD_POLY_PAIRS=${m}u" python fastfourierdiffusion_b200/_build.py >/dev/null 2>&1
  echo "mask ${m}: $(python tools/time_kernels.py 2>&1 | grep attn)"
done
