// tcgen05 / TMEM kernels (sm_100a): the fused FFN + residual + LayerNorm of one encoder layer,
//   y = LN2(x + W2 relu(W1 x + b1) + b2)          cached_transformer.py:130-132,325-327
// which carries 86 % of the sampling path's FLOPs.
//
// One persistent CTA per SM walks over pairs of 128-token tiles (A, B).  Per 128-wide slice j of
// the F hidden units:
//   GEMM1  H_s = X_s[128 x KP] . W1_j^T          tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM)
//   epilogue warps: tcgen05.ld H_s -> relu -> bf16 -> shared (canonical K-major tile P_s)
//   GEMM2  Y_s += P_s[128 x 128] . W2_j^T        accumulates in TMEM over all slices
// so the [tokens x F] hidden activation never leaves the SM.  W1_j / W2_j are pre-packed (ffd_pack_weights)
// as ready-to-use UMMA core-matrix images and arrive by 1-D bulk async copies (cp.async.bulk, mbarrier
// completion) through a 2-stage ring each; both token tiles reuse every weight slice.  b1 rides in two spare
// K columns of W1 (hi + lo bf16 parts against constant-1 columns of X), so the epilogue is relu + convert only.
// The last epilogue adds b2 and the fp32 residual, applies LayerNorm per row (one thread owns a row) and
// stores fp32.
//
// Warp roles (10 warps): 0 = bulk-copy producer, 1 = TMEM allocator + MMA issuer,
// 2..5 = epilogue of tile A, 6..9 = epilogue of tile B.
#include <cuda_bf16.h>

#include "ffd_internal.h"

namespace ffd {

namespace tc {

constexpr int TILE_M = 128;   // rows per token tile (UMMA M)
constexpr int CHUNK = 128;    // hidden units per slice (UMMA N of GEMM1, K extent of GEMM2)
constexpr int STAGES = 2;     // weight ring depth
constexpr int NUM_THREADS = 320;

__host__ __device__ constexpr int dp_of(int d) { return (d + 2 + 15) / 16 * 16; }

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle: 8x(16 B) core matrices; LBO = byte step
// between the two K halves of one MMA (k-groups), SBO = byte step between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

#define FFD_TMEM_LD16(taddr, r)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// relu + round-to-nearest bf16 of two floats in one instruction; lo -> low half (lower k index)
__device__ __forceinline__ uint32_t relu_pack_bf16(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- shared-memory plan ------------------------------------------------------------------------
template <int DP>
struct Smem {
  static constexpr int X_BYTES = TILE_M * DP * 2;     // bf16 [128 x DP]
  static constexpr int P_BYTES = TILE_M * CHUNK * 2;  // bf16 [128 x 128]
  static constexpr int W1_BYTES = CHUNK * DP * 2;     // bf16 [128 n x DP k]
  static constexpr int W2_BYTES = DP * CHUNK * 2;     // bf16 [DP n x 128 k]
  static constexpr int OFF_X = 0;
  static constexpr int OFF_P = OFF_X + 2 * X_BYTES;
  static constexpr int OFF_W1 = OFF_P + 2 * P_BYTES;
  static constexpr int OFF_W2 = OFF_W1 + STAGES * W1_BYTES;
  static constexpr int OFF_BAR = OFF_W2 + STAGES * W2_BYTES;
  static constexpr int NUM_BARS = 4 * STAGES + 12;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
};

// ---- the kernel ------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_ln_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ w1img,
                 const uint8_t* __restrict__ w2img, const float* __restrict__ b2, const float* __restrict__ lnw,
                 const float* __restrict__ lnb, int64_t M, int d, int n_chunks, int64_t n_pairs, float eps) {
  using S = Smem<DP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* w1_full = bars;                 // [STAGES]
  uint64_t* w1_empty = bars + STAGES;       // [STAGES]
  uint64_t* w2_full = bars + 2 * STAGES;    // [STAGES]
  uint64_t* w2_empty = bars + 3 * STAGES;   // [STAGES]
  uint64_t* x_full = bars + 4 * STAGES;     // [2] 128 arrivals (epilogue threads)
  uint64_t* h_full = x_full + 2;            // [2] commit
  uint64_t* p_full = h_full + 2;            // [2] 128 arrivals
  uint64_t* p_empty = p_full + 2;           // [2] commit
  uint64_t* y_full = p_empty + 2;           // [2] commit
  uint64_t* y_empty = y_full + 2;           // [2] 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 128);
      mbar_init(&h_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&p_empty[s], 1);
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: H_A [0,128) H_B [128,256) Y_A [256,256+DP) Y_B [384,384+DP)

  if (warp == 0) {
    // ================= bulk-copy producer =================
    if (lane == 0) {
      uint32_t k1 = 0, k2 = 0;  // loads issued into each ring
      auto load_w1 = [&](int j) {
        const uint32_t st = k1 % STAGES, u = k1 / STAGES;
        mbar_wait(&w1_empty[st], (u & 1) ^ 1);
        mbar_expect_tx(&w1_full[st], S::W1_BYTES);
        bulk_load(smem + S::OFF_W1 + st * S::W1_BYTES, w1img + (size_t)j * S::W1_BYTES, S::W1_BYTES, &w1_full[st]);
        ++k1;
      };
      auto load_w2 = [&](int j) {
        const uint32_t st = k2 % STAGES, u = k2 / STAGES;
        mbar_wait(&w2_empty[st], (u & 1) ^ 1);
        mbar_expect_tx(&w2_full[st], S::W2_BYTES);
        bulk_load(smem + S::OFF_W2 + st * S::W2_BYTES, w2img + (size_t)j * S::W2_BYTES, S::W2_BYTES, &w2_full[st]);
        ++k2;
      };
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        load_w1(0);  // same order as the MMA warp consumes: W1_0, then W1_{j+1}, W2_j
        for (int j = 0; j < n_chunks; ++j) {
          if (j + 1 < n_chunks) load_w1(j + 1);
          load_w2(j);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t IDESC1 = make_idesc(TILE_M, CHUNK);
      constexpr uint32_t IDESC2 = make_idesc(TILE_M, DP);
      constexpr uint32_t LBO_ROWS128 = 128 * 16;  // k-group stride of a 128-row operand
      constexpr uint32_t LBO_ROWSDP = DP * 16;    // k-group stride of a DP-row operand
      uint32_t c1 = 0, c2 = 0;  // ring consumption counters
      uint32_t nchunk = 0;      // chunks processed so far by this CTA (both tiles advance together)
      uint32_t npair = 0;
      const uint32_t xa = smem_u32(smem + S::OFF_X), pa = smem_u32(smem + S::OFF_P);
      auto gemm1 = [&](int s, uint32_t w1addr) {
        const uint32_t xs = xa + s * S::X_BYTES;
#pragma unroll
        for (int k = 0; k < DP / 16; ++k)
          umma(tmem_base + s * 128, make_desc(xs + k * 2 * LBO_ROWS128, LBO_ROWS128, 128),
               make_desc(w1addr + k * 2 * LBO_ROWS128, LBO_ROWS128, 128), IDESC1, k > 0);
      };
      auto gemm2 = [&](int s, uint32_t w2addr, bool first) {
        const uint32_t ps = pa + s * S::P_BYTES;
#pragma unroll
        for (int k = 0; k < CHUNK / 16; ++k)
          umma(tmem_base + 256 + s * 128, make_desc(ps + k * 2 * LBO_ROWS128, LBO_ROWS128, 128),
               make_desc(w2addr + k * 2 * LBO_ROWSDP, LBO_ROWSDP, 128), IDESC2, (!first || k > 0) ? 1u : 0u);
      };
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++npair) {
        // prologue: GEMM1 of slice 0 for both tiles
        {
          const uint32_t st = c1 % STAGES, u = c1 / STAGES;
          mbar_wait(&w1_full[st], u & 1);
          const uint32_t w1addr = smem_u32(smem + S::OFF_W1 + st * S::W1_BYTES);
          for (int s = 0; s < 2; ++s) {
            mbar_wait(&x_full[s], npair & 1);
            tc_fence_after();
            gemm1(s, w1addr);
            umma_commit(&h_full[s]);
          }
          umma_commit(&w1_empty[st]);
          ++c1;
        }
        for (int j = 0; j < n_chunks; ++j, ++nchunk) {
          const bool more = j + 1 < n_chunks;
          uint32_t w1addr = 0, st1 = 0;
          if (more) {
            st1 = c1 % STAGES;
            mbar_wait(&w1_full[st1], (c1 / STAGES) & 1);
            w1addr = smem_u32(smem + S::OFF_W1 + st1 * S::W1_BYTES);
          }
          const uint32_t st2 = c2 % STAGES;
          mbar_wait(&w2_full[st2], (c2 / STAGES) & 1);
          const uint32_t w2addr = smem_u32(smem + S::OFF_W2 + st2 * S::W2_BYTES);
          for (int s = 0; s < 2; ++s) {
            mbar_wait(&p_full[s], nchunk & 1);  // relu(H_s) of slice j is in P_s; H_s is free again
            tc_fence_after();
            if (more) {
              gemm1(s, w1addr);
              umma_commit(&h_full[s]);
            }
            if (j == 0) {
              mbar_wait(&y_empty[s], (npair & 1) ^ 1);  // previous pair's Y_s has been read out
              tc_fence_after();
            }
            gemm2(s, w2addr, j == 0);
            umma_commit(&p_empty[s]);
            if (!more) umma_commit(&y_full[s]);
          }
          if (more) {
            umma_commit(&w1_empty[st1]);
            ++c1;
          }
          umma_commit(&w2_empty[st2]);
          ++c2;
        }
      }
    }
  } else {
    // ================= epilogue warps: tile s = 0 (warps 2..5) or 1 (warps 6..9) =================
    const int s = (warp - 2) >> 2;
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // row inside the 128-token tile
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t* xs = smem + S::OFF_X + s * S::X_BYTES;
    uint8_t* ps = smem + S::OFF_P + s * S::P_BYTES;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    uint32_t nchunk = 0, npair = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++npair) {
      const int64_t grow = (pair * 2 + s) * TILE_M + row;
      const bool valid = grow < M;
      // ---- stage X_s as bf16 K-major core matrices; columns d, d+1 are the constant-1 bias columns ----
      {
        const float* xr = x + grow * d;
#pragma unroll
        for (int g = 0; g < DP / 8; ++g) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = g * 8 + e;
            v[e] = (c < d) ? (valid ? __ldg(xr + c) : 0.f) : ((c == d || c == d + 1) ? 1.f : 0.f);
          }
          uint4 q;
          q.x = pack_bf16(v[0], v[1]);
          q.y = pack_bf16(v[2], v[3]);
          q.z = pack_bf16(v[4], v[5]);
          q.w = pack_bf16(v[6], v[7]);
          *reinterpret_cast<uint4*>(xs + row_off + g * (TILE_M * 16)) = q;
        }
        fence_proxy_async();
        mbar_arrive(&x_full[s]);
      }
      // ---- per slice: H_s (TMEM fp32) -> relu -> bf16 -> P_s (shared) ----
      for (int j = 0; j < n_chunks; ++j, ++nchunk) {
        mbar_wait(&h_full[s], nchunk & 1);
        tc_fence_after();
        mbar_wait(&p_empty[s], (nchunk & 1) ^ 1);  // GEMM2 of the previous slice has finished reading P_s
#pragma unroll
        for (int c0 = 0; c0 < CHUNK; c0 += 32) {
          uint32_t r0[16], r1[16];
          FFD_TMEM_LD16(tmem_base + lane_base + s * 128 + c0, r0);
          FFD_TMEM_LD16(tmem_base + lane_base + s * 128 + c0 + 16, r1);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 q;
            q.x = relu_pack_bf16(r0[g * 8 + 0], r0[g * 8 + 1]);
            q.y = relu_pack_bf16(r0[g * 8 + 2], r0[g * 8 + 3]);
            q.z = relu_pack_bf16(r0[g * 8 + 4], r0[g * 8 + 5]);
            q.w = relu_pack_bf16(r0[g * 8 + 6], r0[g * 8 + 7]);
            *reinterpret_cast<uint4*>(ps + row_off + (c0 / 8 + g) * (TILE_M * 16)) = q;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 q;
            q.x = relu_pack_bf16(r1[g * 8 + 0], r1[g * 8 + 1]);
            q.y = relu_pack_bf16(r1[g * 8 + 2], r1[g * 8 + 3]);
            q.z = relu_pack_bf16(r1[g * 8 + 4], r1[g * 8 + 5]);
            q.w = relu_pack_bf16(r1[g * 8 + 6], r1[g * 8 + 7]);
            *reinterpret_cast<uint4*>(ps + row_off + (c0 / 8 + 2 + g) * (TILE_M * 16)) = q;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&p_full[s]);
      }
      // ---- final: Y_s + b2 + residual -> LayerNorm -> global ----
      mbar_wait(&y_full[s], npair & 1);
      tc_fence_after();
      float v[DP];
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 16) {
        uint32_t r[16];
        FFD_TMEM_LD16(tmem_base + lane_base + 256 + s * 128 + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[c0 + e] = __uint_as_float(r[e]);
      }
      tc_fence_before();
      mbar_arrive(&y_empty[s]);
      if (valid) {
        const float* xr = x + grow * d;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
          if (c < d) {
            v[c] = v[c] + __ldg(b2 + c) + __ldg(xr + c);
            sum += v[c];
          }
        }
        const float mean = sum / (float)d;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
          if (c < d) {
            const float t = v[c] - mean;
            var = fmaf(t, t, var);
          }
        }
        const float rstd = rsqrtf(var / (float)d + eps);
        float* yr = y + grow * d;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
          if (c < d) yr[c] = (v[c] - mean) * rstd * __ldg(lnw + c) + __ldg(lnb + c);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- weight packing ----------------------------------------------------------------------------------
// Per layer: n_chunks W1 images [128 n x DP k] then n_chunks W2 images [DP n x 128 k], bf16, K-major
// core-matrix order: byte(n, k) = (n/8)*128 + (k/8)*(rows*16) + (n%8)*16 + (k%8)*2.
__global__ void pack_ffn_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                const float* __restrict__ W2, __nv_bfloat16* __restrict__ out, int d, int F, int DP) {
  const int n_chunks = F / CHUNK;
  const int64_t per_img = (int64_t)CHUNK * DP;
  const int64_t total = 2 * n_chunks * per_img;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const bool is_w2 = i >= n_chunks * per_img;
    const int64_t r = is_w2 ? i - n_chunks * per_img : i;
    const int j = (int)(r / per_img);
    const int e = (int)(r % per_img);  // element index inside the image, in bf16 units
    const int rows = is_w2 ? DP : CHUNK;
    // invert the layout: e*2 bytes = (n/8)*128 + (k/8)*(rows*16) + (n%8)*16 + (k%8)*2
    const int kg = e / (rows * 8);
    const int rem = e % (rows * 8);
    const int ng = rem / 64, n8 = (rem % 64) / 8, k8 = rem % 8;
    const int n = ng * 8 + n8, k = kg * 8 + k8;
    float v = 0.f;
    if (!is_w2) {
      const int f = j * CHUNK + n;  // hidden unit; k = input feature (or bias column)
      if (k < d) {
        v = W1[(int64_t)f * d + k];
      } else if (k == d) {
        v = __bfloat162float(__float2bfloat16_rn(b1[f]));
      } else if (k == d + 1) {
        v = b1[f] - __bfloat162float(__float2bfloat16_rn(b1[f]));
      }
    } else {
      const int f = j * CHUNK + k;  // n = output feature, k = hidden unit inside the slice
      if (n < d) v = W2[(int64_t)n * F + f];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace tc

bool tc_supported(const ffd_dims& m) { return m.d_model + 2 <= 80 && m.d_ff % tc::CHUNK == 0 && m.d_ff >= tc::CHUNK; }

static int64_t layer_image_bytes(const ffd_dims& m) {
  const int DP = tc::dp_of(m.d_model);
  return (int64_t)2 * (m.d_ff / tc::CHUNK) * tc::CHUNK * DP * 2;
}

int64_t tc_packed_bytes(const ffd_dims& m) {
  if (!tc_supported(m)) return 0;
  return layer_image_bytes(m) * m.n_layers;
}

int tc_pack_weights(const ffd_dims& m, const float* weights, void* packed, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const int DP = tc::dp_of(m.d_model);
  for (int l = 0; l < m.n_layers; ++l) {
    const float* w = weights + ar.layer0 + (int64_t)l * ar.layer_stride;
    tc::pack_ffn_kernel<<<256, 256, 0, s>>>(w + ar.w1, w + ar.b1, w + ar.w2,
                                            (__nv_bfloat16*)((uint8_t*)packed + l * layer_image_bytes(m)),
                                            m.d_model, m.d_ff, DP);
    FFD_CHECK_LAUNCH("pack_ffn_kernel");
  }
  return FFD_OK;
}

template <int DP>
static int launch_ffn_tc(const ffd_dims& m, const uint8_t* img, const float* w, const Arena& ar, const float* x,
                         float* y, int64_t M, cudaStream_t s) {
  using S = tc::Smem<DP>;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    FFD_CHECK_CUDA(cudaGetDevice(&dev));
    FFD_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::ffn_ln_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::TOTAL));
  }
  const int n_chunks = m.d_ff / tc::CHUNK;
  const int64_t n_pairs = (M + 2 * tc::TILE_M - 1) / (2 * tc::TILE_M);
  const unsigned grid = (unsigned)(n_pairs < n_sm ? n_pairs : n_sm);
  const uint8_t* w1img = img;
  const uint8_t* w2img = img + (size_t)n_chunks * S::W1_BYTES;
  tc::ffn_ln_tc_kernel<DP><<<grid, tc::NUM_THREADS, S::TOTAL, s>>>(x, y, w1img, w2img, w + ar.b2, w + ar.n2w,
                                                                    w + ar.n2b, M, m.d_model, n_chunks, n_pairs,
                                                                    1e-5f);
  FFD_CHECK_LAUNCH("ffn_ln_tc_kernel");
  return FFD_OK;
}

int tc_ffn_ln(const ffd_dims& m, const void* packed, const float* weights, int layer, const float* x, float* y,
              int64_t M, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const uint8_t* img = (const uint8_t*)packed + (size_t)layer * layer_image_bytes(m);
  switch (tc::dp_of(m.d_model)) {
    case 16: return launch_ffn_tc<16>(m, img, w, ar, x, y, M, s);
    case 32: return launch_ffn_tc<32>(m, img, w, ar, x, y, M, s);
    case 48: return launch_ffn_tc<48>(m, img, w, ar, x, y, M, s);
    case 64: return launch_ffn_tc<64>(m, img, w, ar, x, y, M, s);
    case 80: return launch_ffn_tc<80>(m, img, w, ar, x, y, M, s);
  }
  set_error("tensor-core FFN: unsupported d_model %d", m.d_model);
  return FFD_ERR_UNSUPPORTED;
}

}  // namespace ffd
