// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for several N, A from shared memory (SS) or TMEM (TS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I fastfourierdiffusion_b200/csrc tools/mma_bench.cu -o /tmp/mma_bench
#include <cstdio>
#include "ffd_tc_ptx.cuh"
namespace ffd { void set_error(const char*, ...) {} void count_launch() {} }
using namespace ffd::tc;

__global__ void __launch_bounds__(128, 1) bench(int N, int ts, int count, int dep, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) smem[i] = 0;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, slot, 0);
  if (warp == 1) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
    const uint64_t ad = make_desc(a, 2048, 128), bd = make_desc(b, N * 16, 128);
    long long t0 = clock64();
    if (dep == 2) {
      // straight-line: 8 MMAs per asm block under one election, no per-MMA overhead
      for (int i = 0; i < count / 8; ++i) {
        if (ts)
          asm volatile(
              "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t}" ::"r"(tb),
              "r"(tb + 480), "l"(bd), "r"(idesc)
              : "memory");
        else
          asm volatile(
              "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t}" ::"r"(tb),
              "l"(ad), "l"(bd), "r"(idesc)
              : "memory");
      }
    } else {
      for (int i = 0; i < count; ++i) {
        const uint32_t dcol = dep ? 0u : (uint32_t)((i & 1) * 256);
        if (ts) umma_ts(tb + dcol, tb + 480, bd, idesc, 1);
        else umma(tb + dcol, ad, bd, idesc, 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int Ns[] = {16, 32, 64, 80, 128, 192, 256};
  printf("mode N count dep issue_cyc/mma total_cyc/mma\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int N : Ns)
      for (int dep = 1; dep < 3; ++dep) {
        const int count = 256;
        bench<<<1, 128, 64 * 1024>>>(N, ts, count, dep, d);
        bench<<<1, 128, 64 * 1024>>>(N, ts, count, dep, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%s %3d %d %d %7.1f %7.1f %s\n", ts ? "TS" : "SS", N, count, dep, (double)h[0] / count, (double)h[1] / count,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
