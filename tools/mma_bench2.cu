// Microbenchmark 2: what bounds the two MMA-issuing warps of the fused layer kernel?
//   attn step: 6 TS MMAs (N = 16) + 2 commits + 1 SS MMA (N = 96) + 2 commits          (ffd_tc_ptx.cuh attn_issue_step)
//   ffn step : 4 TS MMAs (N = 80) + 1 commit + 5 TS MMAs (N = 64) + 2 commits          (ffn_issue_step)
// each alone, both at once from two warps, with / without 16 "softmax" warps doing TMEM loads, MUFU and TMEM stores,
// with the commits removed, and with each warp's tcgen05 instructions split over more elections.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I fastfourierdiffusion_b200/csrc tools/mma_bench2.cu -o tools/mma_bench2.bin
#include <cstdio>
#include "ffd_tc_ptx.cuh"
namespace ffd { void set_error(const char*, ...) {} void count_launch() {} }
using namespace ffd::tc;

// mode bits: 1 attn warp active, 2 ffn warp active, 4 softmax load, 8 no commits, 16 commits only
__global__ void __launch_bounds__(1024, 1) bench(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1u << 20); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 1024) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, slot, 0);
  const uint32_t b0 = smem_u32(&bars[0]);
  const uint32_t sa = smem_u32(smem);
  if (warp == 29 && (mode & 1)) {
    const uint64_t vd = make_desc(sa, 256, 128), qd = make_desc(sa + 8192, 2048, 128), kd = make_desc(sa + 16384, 96 * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (mode & 16) {
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
          "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
          "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t"
          "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
          "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%1];\n\t}" ::"r"(b0), "r"(b0 + 8) : "memory");
      } else if (mode & 8) {
        attn_issue_step(tb + 192, tb, vd, 32, make_idesc_f16(128, 16), b0, b0 + 8, 1, tb, qd, kd, make_idesc(128, 96), b0 + 16, b0 + 24);
      } else {
        attn_issue_step(tb + 192, tb, vd, 32, make_idesc_f16(128, 16), b0, b0 + 8, 1, tb, qd, kd, make_idesc(128, 96), b0 + 16, b0 + 24);
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / iters;
  }
  if (warp == 31 && (mode & 2)) {
    const uint64_t w2 = make_desc(sa + 24576, 80 * 16, 128), w1 = make_desc(sa + 40960, 64 * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
      ffn_issue_step(tb + 304, tb + 384, w2, 160, make_idesc(128, 80), b0 + 32, b0 + 40, 0, 1, tb + 384 + 64 * (i & 1), tb + 256, w1,
                     128, make_idesc(128, 64), b0 + 48, b0 + 56);
    const long long t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / iters;
  }
  if (warp < 16 && (mode & 4)) {
    // softmax-like load: per "head" 3 x LD16, 48 ex2, 3 x ST8 on this warp's lane quadrant
    const uint32_t col = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 48;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r[48];
      FFD_TMEM_LD16(col, r); FFD_TMEM_LD16(col + 16, (r + 16)); FFD_TMEM_LD16(col + 32, (r + 32));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) pk[e >> 1] = ex2_pack_f16_shift(__uint_as_float(r[16 * j + e]), __uint_as_float(r[16 * j + e + 1]), -1.f);
        FFD_TMEM_ST8(col + 8 * j, pk);
        acc += __uint_as_float(pk[0]);
      }
      tmem_st_wait();
    }
    const long long t1 = clock64();
    if (warp == 0 && lane == 0) out[2] = (t1 - t0) / iters;
    if (acc == 123.456f) out[3] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int modes[] = {1, 2, 3, 4, 5, 6, 7, 1 | 16, 1 | 16 | 4};
  printf("mode(1 attn,2 ffn,4 softmax,16 commits only)  attn_step_cyc  ffn_step_cyc  softmax_head_cyc\n");
  for (int m : modes) {
    cudaMemset(d, 0, 64);
    bench<<<1, 1024, 64 * 1024>>>(m, 200, d);
    bench<<<1, 1024, 64 * 1024>>>(m, 200, d);
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("%2d  %6lld %6lld %6lld %s\n", m, h[0], h[1], h[2], e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
