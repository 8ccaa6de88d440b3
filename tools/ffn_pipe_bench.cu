// Microbenchmark 3: the FFN half-slice pipeline of the fused layer kernel on its own (no weights traffic, no attention):
//   issuer warp:   wait r_full[b] -> GEMM2(j) + GEMM1(j+2) + commits          (ffn_issue_step)
//   epilogue warps (one per lane quadrant): wait h_full[b] -> tcgen05.ld 64 cols -> relu/pack -> tcgen05.st -> arrive
// Variants of the waiting style and of the epilogue, cycles per half-slice.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I fastfourierdiffusion_b200/csrc tools/ffn_pipe_bench.cu -o tools/ffn_pipe_bench.bin
#include <cstdio>
#include "ffd_tc_ptx.cuh"
namespace ffd { void set_error(const char*, ...) {} void count_launch() {} }
using namespace ffd::tc;

__device__ __forceinline__ bool try_wait_nohint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// style 0: try_wait with 20 us hint (product), 1: try_wait without hint, 2: test_wait spin
__device__ __forceinline__ void wait_style(int style, uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  if (style == 0) { while (!mbar_try_wait(bar, parity)) {} }
  else if (style == 1) { while (!try_wait_nohint(a, parity)) {} }
  else { while (!test_wait(a, parity)) {} }
}

// mode: style (0..2) + 8 * (epilogue warps per quadrant - 1) + 16 * (no MMAs: handshake skeleton only)
__global__ void __launch_bounds__(1024, 1) bench(int mode, int n_steps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t h_full[2], r_full[2], sink[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int style = mode & 7, epw = ((mode >> 3) & 1) + 1, nomma = (mode >> 4) & 1;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&h_full[b], 1); mbar_init(&r_full[b], 128); }
    for (int i = 0; i < 4; ++i) mbar_init(&sink[i], 1u << 20);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 1024) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, slot, 0);
  const uint32_t sa = smem_u32(smem);
  if (warp == 31) {
    const uint64_t w2 = make_desc(sa, 80 * 16, 128), w1 = make_desc(sa + 16384, 64 * 16, 128);
    const uint32_t i1 = make_idesc(128, 64), i2 = make_idesc(128, 80);
    const long long t0 = clock64();
    // prologue: GEMM1 of half-slices 0 and 1
    for (int b = 0; b < 2; ++b) {
      if (!nomma) umma_ts_chain<5>(tb + 384 + 64 * b, tb + 256, 8, w1, 128, i1, 0);
      umma_commit(&h_full[b]);
    }
    for (int j = 0; j < n_steps; ++j) {
      const uint32_t b = j & 1;
      wait_style(style, &r_full[b], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t more = j + 2 < n_steps ? 1u : 0u;
      if (!nomma)
        ffn_issue_step(tb + 304, tb + 384 + 64 * b, w2, 160, i2, smem_u32(&sink[0]), smem_u32(&sink[1]), 0, more,
                       tb + 384 + 64 * b, tb + 256, w1, 128, i1, smem_u32(&h_full[b]), smem_u32(&sink[2]));
      else if (more) umma_commit(&h_full[b]);
    }
    const long long t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / n_steps;
  }
  if (warp >= 16 && warp < 16 + 4 * epw) {
    // epw == 2: the second group of four warps takes the odd half-slices (buffer 1), the first the even ones
    const int quad = warp & 3, grp = (warp - 16) >> 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    long long busy = 0;
    for (int i = (epw == 2 ? grp : 0); i < n_steps; i += epw) {
      const uint32_t b = i & 1;
      const uint32_t hcol = tb + lane_base + 384 + 64 * b;
      wait_style(style, &h_full[b], (i >> 1) & 1);
      const long long s0 = clock64();
      tc_fence_after();
      uint32_t r[32], pk[16];
      FFD_TMEM_LD16(hcol, r); FFD_TMEM_LD16(hcol + 16, (r + 16));
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = relu_pack_op(r[2 * e], r[2 * e + 1]);
      FFD_TMEM_LD16(hcol + 32, r); FFD_TMEM_LD16(hcol + 48, (r + 16));
      tmem_ld_wait();
      FFD_TMEM_ST16(hcol, pk);
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = relu_pack_op(r[2 * e], r[2 * e + 1]);
      FFD_TMEM_ST16(hcol + 16, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&r_full[b]);
      busy += clock64() - s0;
    }
    if (warp == 16 && lane == 0) out[1] = busy / ((n_steps + epw - 1) / epw);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("mode = wait style (0 try_wait+hint, 1 try_wait, 2 test_wait) + 8 * (two epilogue groups) + 16 * (no MMAs)\n");
  printf("mode  cycles/half-slice  epilogue busy cycles\n");
  for (int m : {0, 1, 2, 8, 9, 10, 16, 17, 18, 24, 25, 26}) {
    cudaMemset(d, 0, 64);
    bench<<<1, 1024, 64 * 1024>>>(m, 256, d);
    bench<<<1, 1024, 64 * 1024>>>(m, 256, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("%2d  %6lld %6lld %s\n", m, h[0], h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
