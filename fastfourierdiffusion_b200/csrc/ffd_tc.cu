// tcgen05 / TMEM kernels (sm_100a): the fused FFN + residual + LayerNorm of one encoder layer,
//   y = LN2(x + W2 relu(W1 x + b1) + b2)          cached_transformer.py:130-132,325-327
// which carries 86 % of the sampling path's FLOPs.
//
// One persistent CTA per SM walks over pairs of 128-token tiles (A, B).  Per 128-wide slice j of the F hidden
// units:
//   GEMM1  H_s = X_s[128 x DP] . W1_j^T          tcgen05.mma kind::f16, A operand (bf16 X_s) read from TMEM,
//                                                 B operand from shared memory, fp32 accumulate in TMEM
//   epilogue warps: tcgen05.ld H_s -> relu -> bf16 -> tcgen05.st, in place over H_s
//   GEMM2  Y_s += relu(H_s)[128 x 128] . W2_j^T   A operand again from TMEM; accumulates over all slices
// so the [tokens x F] hidden activation never leaves the SM and only the weight slices stream through shared
// memory (pre-packed UMMA core-matrix images, one bulk copy each, 2-stage rings, shared by both token tiles).
// b1 rides in two spare K columns of W1 (hi + lo bf16 parts against constant-1 columns of X), so the slice
// epilogue is relu + convert only.
//
// v6: the slice loop (measured 1350 cycles per slice for both tiles, against 1280 of tensor-pipe time) never
// stops at a pair boundary, and shared memory belongs to the tensor core.  "Tile warps" (one group of four per
// tile) read the NEXT pair's token rows straight from global memory into the bf16 TMEM operand as soon as this
// pair's last GEMM1 has read X_s, and turn Y_s + b2 + residual into the LayerNorm output
// while the next pair's slices already run: the fp32 rows wait in shared memory as the residual, are replaced by
// the pre-LN values in one sweep (so that Y_s is released early) and leave as coalesced 16-byte global stores.  The tensor core streams its B
// operands from shared memory at 64 B/clk, half the SM's shared-memory bandwidth: staging the fp32 tiles through
// shared memory with bulk copies, as v5 did, cost 14 % (tile copies) + 6 % (LayerNorm passes) of the kernel
// (same-box ablations, tools/ab.sh).  The 16 epilogue warps do nothing but slices.
//
// Warp roles (28 warps): 0 = weight-slice producer, 1 / 2 = MMA issuers of tile 0 / tile 1 (two independent issue
// streams), 3 = idle, 4..7 / 8..11 = tile warps of tile 0 / 1, 12..27 = four epilogue groups of four warps: group g
// serves tile g & 1 and the (g >> 1)-th half of every 128-column slice.
#include <stdlib.h>

#include "ffd_tc_ptx.cuh"

#ifndef FFD_FFN_CTA2_DEFAULT
#define FFD_FFN_CTA2_DEFAULT 1
#endif

namespace ffd {

namespace tc {

constexpr int CHUNK = 128;    // hidden units per slice (UMMA N of GEMM1, K extent of GEMM2)
constexpr int STAGES = 2;     // weight ring depth (shared memory: two pairs of fp32 row buffers + the rings)
constexpr int NUM_THREADS = 896;

// ---- shared-memory plan ------------------------------------------------------------------------
struct FfnSmem {
  int vbuf, w1, w2, consts, bars, tmem, total;
  int tile_bytes, tile_stride, w1_bytes, w2_bytes;
};

__host__ __device__ inline FfnSmem ffn_smem(int DP, int d) {
  FfnSmem s;
  int o = 0;
  auto take = [&](int bytes) {
    int r = o;
    o += (bytes + 127) / 128 * 128;
    return r;
  };
  s.tile_bytes = TILE_M * d * 4;  // fp32 [128 x d], row-contiguous exactly as in HBM
  s.tile_stride = (s.tile_bytes + 127) / 128 * 128;
  s.w1_bytes = CHUNK * DP * 2;    // bf16 [128 n x DP k]
  s.w2_bytes = DP * CHUNK * 2;    // bf16 [DP n x 128 k]
  s.vbuf = take(4 * s.tile_stride);   // fp32 rows (residual, then pre-LayerNorm values) of tile s of pair parity p: 2 p + s
  s.w1 = take(STAGES * s.w1_bytes);
  s.w2 = take(STAGES * s.w2_bytes);
  s.consts = take(3 * DP * 4);    // b2, ln weight, ln bias
  s.bars = take(40 * 8);
  s.tmem = take(16);
  s.total = o;
  return s;
}

// ---- the kernel ------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_ln_tc_kernel(const float* x, float* y /* may alias x */, const uint8_t* __restrict__ w1img,
                 const uint8_t* __restrict__ w2img, const float* __restrict__ b2, const float* __restrict__ lnw,
                 const float* __restrict__ lnb, int64_t M, int d, int n_chunks, int64_t n_pairs, float eps,
                 long long* __restrict__ dbg) {
#ifdef FFD_FFN_STAMPS  // timeline of the third tile pair of CTA 0 (tools/ffn_stamps.py)
#define FFN_STAMP(cond, slot) do { if (dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define FFN_STAMP(cond, slot) do { } while (0)
#endif
  extern __shared__ __align__(1024) uint8_t smem[];
  const FfnSmem sm = ffn_smem(DP, d);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* w1_full = bars;                 // [STAGES] tx
  uint64_t* w1_empty = bars + STAGES;       // [STAGES] 2 commits (one per issuing warp)
  uint64_t* w2_full = bars + 2 * STAGES;    // [STAGES] tx
  uint64_t* w2_empty = bars + 3 * STAGES;   // [STAGES] 2 commits
  uint64_t* t_full = bars + 4 * STAGES;     // [4] tx: fp32 tile landed in buffer (pair parity, tile)
  uint64_t* y_ready = t_full + 4;           // [4] 128 arrivals: the tile has left the buffer (LayerNorm output stored)
  uint64_t* x_full = y_ready + 4;           // [2] 128 arrivals: bf16 X_s written to TMEM
  uint64_t* x_free = x_full + 2;            // [2] commit: the pair's last GEMM1 has read X_s
  uint64_t* h_full = x_free + 2;            // [2] commit
  uint64_t* p_full = h_full + 2;            // [2] 256 arrivals (both column halves)
  uint64_t* y_full = p_full + 2;            // [2] commit
  uint64_t* y_empty = y_full + 2;           // [2] 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.tmem);
  float* consts = reinterpret_cast<float*>(smem + sm.consts);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int XC = (DP / 2 + 15) / 16 * 16;       // TMEM columns of the packed bf16 X operand
  constexpr uint32_t COL_H = 0, COL_Y = 256, COL_X = 256 + DP;  // + 128 * s
  static_assert(256 + DP + XC <= 384, "TMEM plan");
  const uint32_t my_pairs = (uint32_t)((n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x);  // >= 1

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 2);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&y_ready[i], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 128);
      mbar_init(&x_free[s], 1);
      mbar_init(&h_full[s], 1);
      mbar_init(&p_full[s], 256);
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 128);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 3 * DP; i += NUM_THREADS) {
    const int w = i / DP, c = i % DP;
    consts[i] = c < d ? (w == 0 ? b2[c] : w == 1 ? lnw[c] : lnb[c]) : 0.f;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // provably warp-uniform

  // register budget per warpgroup (896 threads x 72 at launch); set at the top of each role branch so that ptxas
  // allocates per branch
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ================= bulk-copy producer =================
      if (lane == 0) {
        uint32_t k1 = 0, k2 = 0;
        auto load_w1 = [&](int j) {
          const uint32_t st = k1 % STAGES, u = k1 / STAGES;
          mbar_wait(&w1_empty[st], (u & 1) ^ 1);
          mbar_expect_tx(&w1_full[st], sm.w1_bytes);
          bulk_load(smem + sm.w1 + st * sm.w1_bytes, w1img + (size_t)j * sm.w1_bytes, sm.w1_bytes, &w1_full[st]);
          ++k1;
        };
        auto load_w2 = [&](int j) {
          const uint32_t st = k2 % STAGES, u = k2 / STAGES;
          mbar_wait(&w2_empty[st], (u & 1) ^ 1);
          mbar_expect_tx(&w2_full[st], sm.w2_bytes);
          bulk_load(smem + sm.w2 + st * sm.w2_bytes, w2img + (size_t)j * sm.w2_bytes, sm.w2_bytes, &w2_full[st]);
          ++k2;
        };
        for (uint32_t np = 0; np < my_pairs; ++np) {
          // same order as the MMA warps consume: W1_0, then W2_j, W1_{j+1}
          load_w1(0);
          for (int j = 0; j < n_chunks; ++j) {
            load_w2(j);
            if (j + 1 < n_chunks) load_w1(j + 1);
          }
        }
      }
    } else if (warp == 3) {
      // ================= token tiles in (one bulk copy per tile; a tile of 128 rows is contiguous) =================
      if (lane == 0) {
        auto load_pair = [&](uint32_t np) {
          const int64_t pair = (int64_t)blockIdx.x + (int64_t)np * gridDim.x;
          for (int s = 0; s < 2; ++s) {
            const int buf = (np & 1) * 2 + s;
            const int64_t row0 = (pair * 2 + s) * TILE_M;
            const int64_t rows = M - row0 < TILE_M ? (M - row0 > 0 ? M - row0 : 0) : TILE_M;
            if (np >= 2) mbar_wait_relaxed(&y_ready[buf], ((np - 2) >> 1) & 1, 100);  // its previous tile has left
            if (rows > 0) {
              mbar_expect_tx(&t_full[buf], (uint32_t)(rows * d * 4));
              bulk_load(smem + sm.vbuf + buf * sm.tile_stride, x + row0 * d, (uint32_t)(rows * d * 4), &t_full[buf]);
            } else {
              mbar_arrive(&t_full[buf]);
            }
          }
        };
        for (uint32_t np = 0; np < my_pairs; ++np) load_pair(np);
      }
    } else if (warp == 1 || warp == 2) {
      // ================= MMA issuers: warp 1 drives tile 0 of every pair, warp 2 tile 1 (whole warp converged,
      // one elected lane issues).  Two independent issue streams: neither tile's chain waits behind the other
      // tile's barrier, and the per-instruction issue cost (~45 cycles per tcgen05.mma) is paid in parallel. =====
      constexpr uint32_t IDESC1 = make_idesc(TILE_M, CHUNK);
      constexpr uint32_t IDESC2 = make_idesc(TILE_M, DP);
      const int s = warp - 1;
      uint32_t c1 = 0, c2 = 0;  // ring consumption counters
      uint32_t nchunk = 0;      // slices processed so far by this CTA
      auto gemm1 = [&](uint32_t w1addr) {
        umma_ts_chain<DP / 16>(tmem_base + COL_H + s * 128, tmem_base + COL_X + s * 128, 8,
                               make_desc(w1addr, CHUNK * 16, 128), (2 * CHUNK * 16) >> 4, IDESC1, 0);
      };
      auto gemm2 = [&](uint32_t w2addr, bool first) {
        // relu(H) is packed per column half: K-steps 0..3 at H_s + 0.., K-steps 4..7 at H_s + 64..
        const uint64_t bd = make_desc(w2addr, DP * 16, 128);
        umma_ts_chain<4>(tmem_base + COL_Y + s * 128, tmem_base + COL_H + s * 128, 8, bd, (2 * DP * 16) >> 4, IDESC2,
                         first ? 0u : 1u);
        umma_ts_chain<4>(tmem_base + COL_Y + s * 128, tmem_base + COL_H + s * 128 + 64, 8,
                         bd + 4 * (uint64_t)((2 * DP * 16) >> 4), (2 * DP * 16) >> 4, IDESC2, 1u);
      };
      for (uint32_t npair = 0; npair < my_pairs; ++npair) {
        // prologue: GEMM1 of slice 0 (its H columns are free: the previous pair's last GEMM2 precedes it in
        // issue order)
        {
          const uint32_t st = c1 % STAGES, u = c1 / STAGES;
          mbar_wait(&w1_full[st], u & 1);
          const uint32_t w1addr = smem_u32(smem + sm.w1 + st * sm.w1_bytes);
          mbar_wait(&x_full[s], npair & 1);
          tc_fence_after();
          gemm1(w1addr);
          umma_commit(&h_full[s]);
          if (n_chunks == 1) umma_commit(&x_free[s]);
          umma_commit(&w1_empty[st]);
          ++c1;
        }
        for (int j = 0; j < n_chunks; ++j, ++nchunk) {
          const bool more = j + 1 < n_chunks;
          const uint32_t st2 = c2 % STAGES;
          mbar_wait(&w2_full[st2], (c2 / STAGES) & 1);
          const uint32_t w2addr = smem_u32(smem + sm.w2 + st2 * sm.w2_bytes);
          uint32_t w1addr = 0, st1 = 0;
          if (more) {
            st1 = c1 % STAGES;
            mbar_wait(&w1_full[st1], (c1 / STAGES) & 1);
            w1addr = smem_u32(smem + sm.w1 + st1 * sm.w1_bytes);
          }
          mbar_wait(&p_full[s], nchunk & 1);  // relu(H_s) of slice j sits (bf16) in the H_s columns
          FFN_STAMP(npair == 2 && s == 0, j * 8 + 0);
          if (j == 0) mbar_wait(&y_empty[s], (npair & 1) ^ 1);  // previous pair's Y_s has been read out
          tc_fence_after();
          gemm2(w2addr, j == 0);
          FFN_STAMP(npair == 2 && s == 0, j * 8 + 1);
          if (!more) umma_commit(&y_full[s]);
          if (more) {
            gemm1(w1addr);  // issued after GEMM2: it overwrites the columns GEMM2 reads
            umma_commit(&h_full[s]);
            if (j + 2 == n_chunks) umma_commit(&x_free[s]);  // that was the pair's last GEMM1
            FFN_STAMP(npair == 2 && s == 0, j * 8 + 2);
          }
          umma_commit(&w2_empty[st2]);
          ++c2;
          if (more) {
            umma_commit(&w1_empty[st1]);
            ++c1;
          }
        }
      }
      mbar_wait(&y_full[s], (my_pairs - 1) & 1);  // drain
    }
  } else if (warp < 12) {
    // ================= tile warps (group s = tile s of every pair): stage the next pair's X, finish this pair's
    // LayerNorm =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int s = (warp - 4) >> 2;
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // row inside the 128-token tile
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int n4 = d / 4;                              // float4 per row
    // bf16 X_s -> TMEM (A operand of GEMM1), rows read straight from global memory; columns d, d+1 are the
    // constant-1 bias columns
    auto stage_x = [&](uint32_t np) {
      const int64_t pair = (int64_t)blockIdx.x + (int64_t)np * gridDim.x;
      const int64_t grow = (pair * 2 + s) * TILE_M + row;
      const bool valid = grow < M;
      const int buf = (np & 1) * 2 + s;
      const float4* src = reinterpret_cast<const float4*>(smem + sm.vbuf + buf * sm.tile_stride) + (size_t)row * n4;
      mbar_wait_relaxed(&t_full[buf], (np >> 1) & 1, 100);
      // convert first (registers), so that only the TMEM stores remain once X_s is free
      uint32_t xp[XC];
#pragma unroll
      for (int i = 0; i < XC / 2; ++i) {  // 4 columns -> 2 packed words
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = i * 4;
        if (c < d) {
          if (valid) v = src[i];
        } else if (c == d) {
          v.x = 1.f, v.y = 1.f;
        }
        xp[2 * i] = pack_op(v.x, v.y);
        xp[2 * i + 1] = pack_op(v.z, v.w);
      }
      if (np > 0) {
        mbar_wait_relaxed(&x_free[s], (np - 1) & 1, 20);  // the previous pair's last GEMM1 has read X_s
        tc_fence_after();
      }
#pragma unroll
      for (int c0 = 0; c0 < XC; c0 += 16) FFD_TMEM_ST16(tmem_base + lane_base + COL_X + s * 128 + c0, (xp + c0));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&x_full[s]);
    };
    // Y_s + b2 + residual -> LayerNorm -> global.  Sweep 1 (one thread owns a row): pre-LN values into the
    // shared-memory buffer, row statistics; Y_s is released.  Sweep 2: the tile leaves as consecutive 16-byte
    // chunks (thread t takes chunks t, t + 128, ...: coalesced stores).
    auto finish = [&](uint32_t np) {
      float4* vb = reinterpret_cast<float4*>(smem + sm.vbuf + ((np & 1) * 2 + s) * sm.tile_stride);
      const int64_t pair = (int64_t)blockIdx.x + (int64_t)np * gridDim.x;
      const int64_t row0 = (pair * 2 + s) * TILE_M;
      const int rows = (int)(M - row0 < TILE_M ? (M - row0 > 0 ? M - row0 : 0) : TILE_M);
      float4* myrow = vb + (size_t)row * n4;
      mbar_wait_relaxed(&y_full[s], np & 1, 40);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      // 16 accumulator columns + b2 + residual -> the buffer, row statistics
      auto unit = [&](const uint32_t (&r)[16], const float4 (&res)[4], int c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + 4 * i;
          if (c < d) {
            float4 v;
            v.x = __uint_as_float(r[4 * i]) + consts[c] + res[i].x;
            v.y = __uint_as_float(r[4 * i + 1]) + consts[c + 1] + res[i].y;
            v.z = __uint_as_float(r[4 * i + 2]) + consts[c + 2] + res[i].z;
            v.w = __uint_as_float(r[4 * i + 3]) + consts[c + 3] + res[i].w;
            myrow[c >> 2] = v;
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
          }
        }
      };
      auto residual = [&](float4 (&e)[4], int c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + 4 * i;
          e[i] = c < d ? myrow[c >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // two 16-column units per TMEM round trip; the residual loads are issued before the round trip completes
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 32) {
        uint32_t ra[16], rb[16];
        float4 e[4];
        constexpr bool kTwoAlways = (DP % 32) == 0;
        const bool two = kTwoAlways || c0 + 16 < DP;
        FFD_TMEM_LD16(tmem_base + lane_base + COL_Y + s * 128 + c0, ra);
        if (two) FFD_TMEM_LD16(tmem_base + lane_base + COL_Y + s * 128 + c0 + 16, rb);
        residual(e, c0);
        tmem_ld_wait();
        if (c0 + 32 >= DP) {
          tc_fence_before();
          mbar_arrive(&y_empty[s]);
        }
        unit(ra, e, c0);
        if (two) {
          residual(e, c0 + 16);
          unit(rb, e, c0 + 16);
        }
      }
      const float mean = s1 / (float)d;
      const float rstd = rsqrtf(fmaxf(s2 / (float)d - mean * mean, 0.f) + eps);
      __syncwarp();  // this warp's 32 rows are in the buffer
      {
        // sweep 2: the warp's 32 rows are one contiguous block of the tile; lane l takes its 16-byte chunks l,
        // l + 32, ... (coalesced global stores); the row statistics come from the lane that owns the row
        const int wrow0 = quad * 32;
        const float4* wv = vb + (size_t)wrow0 * n4;
        float4* dst = reinterpret_cast<float4*>(y + (row0 + wrow0) * d);
        const int wrows = rows - wrow0 < 32 ? rows - wrow0 : 32;  // may be <= 0 in the ragged last tile
        const int step_r = 32 / n4, step_c = 32 % n4;
        int r = lane / n4, c4 = lane % n4;
        for (int q = lane; q < 32 * n4; q += 32) {
          const float m_r = __shfl_sync(0xffffffffu, mean, r & 31);
          const float s_r = __shfl_sync(0xffffffffu, rstd, r & 31);
          if (r < wrows) {
            const float4 v = wv[q];
            const float4 w4 = *reinterpret_cast<const float4*>(consts + DP + 4 * c4);
            const float4 b4 = *reinterpret_cast<const float4*>(consts + 2 * DP + 4 * c4);
            float4 o;
            o.x = (v.x - m_r) * s_r * w4.x + b4.x;
            o.y = (v.y - m_r) * s_r * w4.y + b4.y;
            o.z = (v.z - m_r) * s_r * w4.z + b4.z;
            o.w = (v.w - m_r) * s_r * w4.w + b4.w;
            if (eps < 0.f) o = v;  // debug: pre-LayerNorm values (FFD_FFN_RAW)
            dst[q] = o;
          }
          r += step_r, c4 += step_c;
          if (c4 >= n4) c4 -= n4, ++r;
        }
      }
      fence_proxy_async();  // the next bulk copy into this buffer is ordered after these reads
      mbar_arrive(&y_ready[(np & 1) * 2 + s]);
    };
    stage_x(0);
    for (uint32_t np = 0; np < my_pairs; ++np) {
      FFN_STAMP(warp == 4 && np == 2, 200);
      if (np + 1 < my_pairs) stage_x(np + 1);
      FFN_STAMP(warp == 4 && np == 2, 201);
      finish(np);
      FFN_STAMP(warp == 4 && np == 2, 202);
    }
  } else {
    // ================= epilogue groups (slices only): tile s = g & 1, column half hf = g >> 1 =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    const int g = (warp - 12) >> 2;
    const int s = g & 1, hf = g >> 1;
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t hcol = tmem_base + lane_base + COL_H + s * 128 + hf * 64;  // this group's 64 fp32 columns
    const uint32_t n_slices = my_pairs * (uint32_t)n_chunks;
    // this group's half of H_s (TMEM fp32) -> relu -> bf16, written back in place
    for (uint32_t nchunk = 0; nchunk < n_slices; ++nchunk) {
      mbar_wait_relaxed(&h_full[s], nchunk & 1, 20);
      tc_fence_after();
      FFN_STAMP(warp == 12 && nchunk - 2u * n_chunks < (uint32_t)n_chunks, (nchunk - 2u * n_chunks) * 8 + 3);
      uint32_t r[64], pk[32];
      FFD_TMEM_LD16(hcol, r);
      FFD_TMEM_LD16(hcol + 16, (r + 16));
      FFD_TMEM_LD16(hcol + 32, (r + 32));
      FFD_TMEM_LD16(hcol + 48, (r + 48));
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) pk[e] = relu_pack_op(r[2 * e], r[2 * e + 1]);
      // 64 fp32 columns -> 32 packed columns at the start of this group's own half
      FFD_TMEM_ST16(hcol, pk);
      FFD_TMEM_ST16(hcol + 16, (pk + 16));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[s]);
      FFN_STAMP(warp == 12 && nchunk - 2u * n_chunks < (uint32_t)n_chunks, (nchunk - 2u * n_chunks) * 8 + 4);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

constexpr int STAGES2 = 4;    // weight ring depth of the CTA-pair kernel (half slices: 10 KB each at DP = 80)

__host__ __device__ inline FfnSmem ffn2_smem(int DP, int d) {
  FfnSmem s;
  int o = 0;
  auto take = [&](int bytes) {
    int r = o;
    o += (bytes + 127) / 128 * 128;
    return r;
  };
  s.tile_bytes = TILE_M * d * 4;
  s.tile_stride = (s.tile_bytes + 127) / 128 * 128;
  s.w1_bytes = (CHUNK / 2) * DP * 2;    // bf16 [64 n x DP k]: this CTA's half of a W1 slice
  s.w2_bytes = (DP / 2) * CHUNK * 2;    // bf16 [DP/2 n x 128 k]: this CTA's half of a W2 slice
  s.vbuf = take(4 * s.tile_stride);
  s.w1 = take(STAGES2 * s.w1_bytes);
  s.w2 = take(STAGES2 * s.w2_bytes);
  s.consts = take(3 * DP * 4);
  s.bars = take(64 * 8);
  s.tmem = take(16);
  s.total = o;
  return s;
}

// ---- the same kernel on CTA PAIRS (cta_group::2) -------------------------------------------------------
// Two CTAs of a cluster (two SMs of one TPC) run every MMA together as one M = 256 instruction: CTA r supplies its
// own tile as the A rows (TMEM) and only HALF of the B operand (N / 2 rows of the weight slice) from its shared
// memory.  Per SM this halves the tensor core's shared-memory operand traffic (64 -> 32 B/clk) and the weight-slice
// stream -- shared memory bandwidth is what bounds the one-CTA kernel (see its header).  The leader CTA (rank 0)
// issues all MMAs; every barrier an issuing warp waits on lives in the leader and is arrived on by both CTAs
// (remote arrives through mapa), every completion is committed to both CTAs (multicast commit).  Weight slices are
// packed a second time as per-CTA halves (pack_ffn_halves_kernel).  Everything else -- tile warps, epilogue warps,
// tile loader -- is the one-CTA code, per CTA.
template <int DP, int D_MODEL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
ffn_ln_tc2_kernel(const float* x, float* y /* may alias x */, const uint8_t* __restrict__ w1img,
                 const uint8_t* __restrict__ w2img, const float* __restrict__ b2, const float* __restrict__ lnw,
                 const float* __restrict__ lnb, int64_t M, int /* d = D_MODEL */, int n_chunks, int64_t n_pairs, float eps,
                 long long* __restrict__ dbg) {
#ifdef FFD_FFN_STAMPS  // timeline of the third tile pair of CTA 0 (tools/ffn_stamps.py)
#define FFN_STAMP(cond, slot) do { if (dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define FFN_STAMP(cond, slot) do { } while (0)
#endif
  // d_model is a compile-time constant here: with a run-time d the unrolled X staging (which column holds the
  // constant-1 bias pair) compiled to spilling code whose second inlined copy lost the bias-low column
  constexpr int d = D_MODEL;
  static_assert(D_MODEL % 4 == 0 && D_MODEL + 2 <= DP && D_MODEL + 16 > DP, "d_model / DP");
  extern __shared__ __align__(1024) uint8_t smem[];
  const FfnSmem sm = ffn2_smem(DP, d);
  const uint32_t rank = cluster_ctarank();   // 0 = leader: issues the MMAs of the pair
  const bool leader = rank == 0;
  const uint32_t n_clu = gridDim.x >> 1, clu = blockIdx.x >> 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  constexpr int STAGES = STAGES2;
  uint64_t* w1_full = bars;                 // [STAGES] tx: this CTA's half of a W1 slice landed
  uint64_t* w1_empty = bars + STAGES;       // [STAGES] 2 commits (one per issuing warp of the leader, multicast)
  uint64_t* w2_full = bars + 2 * STAGES;    // [STAGES] tx
  uint64_t* w2_empty = bars + 3 * STAGES;   // [STAGES] 2 commits
  uint64_t* pw1_ready = bars + 4 * STAGES;  // [STAGES] leader: 1 arrival from the peer: its half of W1 landed
  uint64_t* pw2_ready = bars + 5 * STAGES;  // [STAGES] leader: ... of W2
  uint64_t* t_full = bars + 6 * STAGES;     // [4] tx: fp32 tile landed in buffer (pair parity, tile)
  uint64_t* y_ready = t_full + 4;           // [4] 128 arrivals: the tile has left the buffer (LayerNorm output stored)
  uint64_t* x_full = y_ready + 4;           // [2] 128 arrivals: bf16 X_s written to TMEM
  uint64_t* x_free = x_full + 2;            // [2] commit: the pair's last GEMM1 has read X_s
  uint64_t* h_full = x_free + 2;            // [2] commit
  uint64_t* p_full = h_full + 2;            // [2] 256 arrivals (both column halves)
  uint64_t* y_full = p_full + 2;            // [2] commit
  uint64_t* y_empty = y_full + 2;           // [2] 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.tmem);
  float* consts = reinterpret_cast<float*>(smem + sm.consts);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int XC = (DP / 2 + 15) / 16 * 16;       // TMEM columns of the packed bf16 X operand
  constexpr uint32_t COL_H = 0, COL_Y = 256, COL_X = 256 + DP;  // + 128 * s
  static_assert(256 + DP + XC <= 384, "TMEM plan");
  // A cluster walks over QUADS of tiles: CTA r owns pair 2 quad + r (tiles 4 quad + 2 r + s); both CTAs run the same
  // number of iterations (pairs beyond the end have no rows).
  const int64_t n_quads = (n_pairs + 1) / 2;
  const uint32_t my_pairs = (uint32_t)((n_quads - clu + n_clu - 1) / n_clu);  // >= 1
  auto pair_of = [&](uint32_t np) { return ((int64_t)clu + (int64_t)np * n_clu) * 2 + rank; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 2);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 2);
      mbar_init(&pw1_ready[i], 1);
      mbar_init(&pw2_ready[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&y_ready[i], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 256);   // both CTAs (leader only)
      mbar_init(&x_free[s], 1);
      mbar_init(&h_full[s], 1);
      mbar_init(&p_full[s], 512);
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 256);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 3 * DP; i += NUM_THREADS) {
    const int w = i / DP, c = i % DP;
    consts[i] = c < d ? (w == 0 ? b2[c] : w == 1 ? lnw[c] : lnb[c]) : 0.f;
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // provably warp-uniform

  // register budget per warpgroup (896 threads x 72 at launch); set at the top of each role branch so that ptxas
  // allocates per branch
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ================= bulk-copy producer =================
      if (lane == 0) {
        uint32_t k1 = 0, k2 = 0;
        auto load_w1 = [&](int j) {
          const uint32_t st = k1 % STAGES, u = k1 / STAGES;
          mbar_wait(&w1_empty[st], (u & 1) ^ 1);
          mbar_expect_tx(&w1_full[st], sm.w1_bytes);
          bulk_load(smem + sm.w1 + st * sm.w1_bytes, w1img + ((size_t)j * 2 + rank) * sm.w1_bytes, sm.w1_bytes, &w1_full[st]);
          ++k1;
        };
        auto load_w2 = [&](int j) {
          const uint32_t st = k2 % STAGES, u = k2 / STAGES;
          mbar_wait(&w2_empty[st], (u & 1) ^ 1);
          mbar_expect_tx(&w2_full[st], sm.w2_bytes);
          bulk_load(smem + sm.w2 + st * sm.w2_bytes, w2img + ((size_t)j * 2 + rank) * sm.w2_bytes, sm.w2_bytes, &w2_full[st]);
          ++k2;
        };
        for (uint32_t np = 0; np < my_pairs; ++np) {
          // same order as the MMA warps consume: W1_0, then W2_j, W1_{j+1}
          load_w1(0);
          for (int j = 0; j < n_chunks; ++j) {
            load_w2(j);
            if (j + 1 < n_chunks) load_w1(j + 1);
          }
        }
      }
    } else if (warp == 3) {
      // ================= token tiles in (one bulk copy per tile; a tile of 128 rows is contiguous) =================
      if (lane == 0) {
        auto load_pair = [&](uint32_t np) {
          const int64_t pair = pair_of(np);
          for (int s = 0; s < 2; ++s) {
            const int buf = (np & 1) * 2 + s;
            const int64_t row0 = (pair * 2 + s) * TILE_M;
            const int64_t rows = M - row0 < TILE_M ? (M - row0 > 0 ? M - row0 : 0) : TILE_M;
            if (np >= 2) mbar_wait_relaxed(&y_ready[buf], ((np - 2) >> 1) & 1, 100);  // its previous tile has left
            if (rows > 0) {
              mbar_expect_tx(&t_full[buf], (uint32_t)(rows * d * 4));
              bulk_load(smem + sm.vbuf + buf * sm.tile_stride, x + row0 * d, (uint32_t)(rows * d * 4), &t_full[buf]);
            } else {
              mbar_arrive(&t_full[buf]);
            }
          }
        };
        for (uint32_t np = 0; np < my_pairs; ++np) load_pair(np);
      }
    } else if (leader) {
      // ================= MMA issuers of the PAIR (leader CTA): warp 1 drives tile 0 of both CTAs, warp 2 tile 1
      // (whole warp converged, one elected lane issues) =================
      constexpr uint32_t IDESC1 = make_idesc(2 * TILE_M, CHUNK);
      constexpr uint32_t IDESC2 = make_idesc(2 * TILE_M, DP);
      const int s = warp - 1;
      uint32_t c1 = 0, c2 = 0;  // ring consumption counters
      uint32_t nchunk = 0;      // slices processed so far by this CTA
      auto gemm1 = [&](uint32_t w1addr) {  // B: this CTA's 64 of the slice's 128 rows (the peer holds the others)
        umma2_ts_chain<DP / 16>(tmem_base + COL_H + s * 128, tmem_base + COL_X + s * 128, 8,
                                make_desc(w1addr, (CHUNK / 2) * 16, 128), (2 * (CHUNK / 2) * 16) >> 4, IDESC1, 0);
      };
      auto gemm2 = [&](uint32_t w2addr, bool first) {
        // relu(H) is packed per column half: K-steps 0..3 at H_s + 0.., K-steps 4..7 at H_s + 64..
        const uint64_t bd = make_desc(w2addr, (DP / 2) * 16, 128);
        umma2_ts_chain<4>(tmem_base + COL_Y + s * 128, tmem_base + COL_H + s * 128, 8, bd, (2 * (DP / 2) * 16) >> 4,
                          IDESC2, first ? 0u : 1u);
        umma2_ts_chain<4>(tmem_base + COL_Y + s * 128, tmem_base + COL_H + s * 128 + 64, 8,
                          bd + 4 * (uint64_t)((2 * (DP / 2) * 16) >> 4), (2 * (DP / 2) * 16) >> 4, IDESC2, 1u);
      };
      for (uint32_t npair = 0; npair < my_pairs; ++npair) {
        // prologue: GEMM1 of slice 0
        {
          const uint32_t st = c1 % STAGES, u = c1 / STAGES;
          mbar_wait(&w1_full[st], u & 1);
          mbar_wait_cluster(&pw1_ready[st], u & 1);
          const uint32_t w1addr = smem_u32(smem + sm.w1 + st * sm.w1_bytes);
          mbar_wait_cluster(&x_full[s], npair & 1);
          tc_fence_after();
          gemm1(w1addr);
          umma2_commit(&h_full[s]);
          if (n_chunks == 1) umma2_commit(&x_free[s]);
          umma2_commit(&w1_empty[st]);
          ++c1;
        }
        for (int j = 0; j < n_chunks; ++j, ++nchunk) {
          const bool more = j + 1 < n_chunks;
          const uint32_t st2 = c2 % STAGES;
          mbar_wait(&w2_full[st2], (c2 / STAGES) & 1);
          mbar_wait_cluster(&pw2_ready[st2], (c2 / STAGES) & 1);
          const uint32_t w2addr = smem_u32(smem + sm.w2 + st2 * sm.w2_bytes);
          uint32_t w1addr = 0, st1 = 0;
          if (more) {
            st1 = c1 % STAGES;
            mbar_wait(&w1_full[st1], (c1 / STAGES) & 1);
            mbar_wait_cluster(&pw1_ready[st1], (c1 / STAGES) & 1);
            w1addr = smem_u32(smem + sm.w1 + st1 * sm.w1_bytes);
          }
          mbar_wait_cluster(&p_full[s], nchunk & 1);  // relu(H_s) of slice j sits (bf16) in both CTAs' H_s columns
          if (j == 0) mbar_wait_cluster(&y_empty[s], (npair & 1) ^ 1);  // previous pairs' Y_s have been read out
          tc_fence_after();
          gemm2(w2addr, j == 0);
          if (!more) umma2_commit(&y_full[s]);
          if (more) {
            gemm1(w1addr);  // issued after GEMM2: it overwrites the columns GEMM2 reads
            umma2_commit(&h_full[s]);
            if (j + 2 == n_chunks) umma2_commit(&x_free[s]);  // that was the pair's last GEMM1
          }
          umma2_commit(&w2_empty[st2]);
          ++c2;
          if (more) {
            umma2_commit(&w1_empty[st1]);
            ++c1;
          }
        }
      }
      mbar_wait(&y_full[s], (my_pairs - 1) & 1);  // drain
    } else if (warp == 1) {
      // ================= peer CTA: tell the leader when this CTA's halves of the weight slices have landed, in the
      // order the leader consumes them =================
      if (lane == 0) {
        uint32_t c1 = 0, c2 = 0;
        auto relay1 = [&]() {
          const uint32_t st = c1 % STAGES;
          mbar_wait_relaxed(&w1_full[st], (c1 / STAGES) & 1, 20);
          mbar_arrive_cta0(&pw1_ready[st]);
          ++c1;
        };
        auto relay2 = [&]() {
          const uint32_t st = c2 % STAGES;
          mbar_wait_relaxed(&w2_full[st], (c2 / STAGES) & 1, 20);
          mbar_arrive_cta0(&pw2_ready[st]);
          ++c2;
        };
        for (uint32_t np = 0; np < my_pairs; ++np) {
          relay1();
          for (int j = 0; j < n_chunks; ++j) {
            relay2();
            if (j + 1 < n_chunks) relay1();
          }
        }
      }
    }
  } else if (warp < 12) {
    // ================= tile warps (group s = tile s of every pair): stage the next pair's X, finish this pair's
    // LayerNorm =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int s = (warp - 4) >> 2;
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // row inside the 128-token tile
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int n4 = d / 4;                              // float4 per row
    // bf16 X_s -> TMEM (A operand of GEMM1), rows read straight from global memory; columns d, d+1 are the
    // constant-1 bias columns
    auto stage_x = [&](uint32_t np) {
      const int64_t pair = pair_of(np);
      const int64_t grow = (pair * 2 + s) * TILE_M + row;
      const bool valid = grow < M;
      const int buf = (np & 1) * 2 + s;
      const float4* src = reinterpret_cast<const float4*>(smem + sm.vbuf + buf * sm.tile_stride) + (size_t)row * n4;
      mbar_wait_relaxed(&t_full[buf], (np >> 1) & 1, 100);
      // convert first (registers), so that only the TMEM stores remain once X_s is free
      uint32_t xp[XC];
#pragma unroll
      for (int i = 0; i < XC / 2; ++i) {  // 4 columns -> 2 packed words
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = i * 4;
        if (c < d) {
          if (valid) v = src[i];
        } else if (c == d) {
          v.x = 1.f, v.y = 1.f;
        }
        xp[2 * i] = pack_op(v.x, v.y);
        xp[2 * i + 1] = pack_op(v.z, v.w);
      }
      if (np > 0) {
        mbar_wait_relaxed(&x_free[s], (np - 1) & 1, 20);  // the previous pair's last GEMM1 has read X_s
        tc_fence_after();
      }
#pragma unroll
      for (int c0 = 0; c0 < XC; c0 += 16) FFD_TMEM_ST16(tmem_base + lane_base + COL_X + s * 128 + c0, (xp + c0));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_cta0(&x_full[s]);
    };
    // Y_s + b2 + residual -> LayerNorm -> global.  Sweep 1 (one thread owns a row): pre-LN values into the
    // shared-memory buffer, row statistics; Y_s is released.  Sweep 2: the tile leaves as consecutive 16-byte
    // chunks (thread t takes chunks t, t + 128, ...: coalesced stores).
    auto finish = [&](uint32_t np) {
      float4* vb = reinterpret_cast<float4*>(smem + sm.vbuf + ((np & 1) * 2 + s) * sm.tile_stride);
      const int64_t pair = pair_of(np);
      const int64_t row0 = (pair * 2 + s) * TILE_M;
      const int rows = (int)(M - row0 < TILE_M ? (M - row0 > 0 ? M - row0 : 0) : TILE_M);
      float4* myrow = vb + (size_t)row * n4;
      mbar_wait_relaxed(&y_full[s], np & 1, 40);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      // 16 accumulator columns + b2 + residual -> the buffer, row statistics
      auto unit = [&](const uint32_t (&r)[16], const float4 (&res)[4], int c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + 4 * i;
          if (c < d) {
            float4 v;
            v.x = __uint_as_float(r[4 * i]) + consts[c] + res[i].x;
            v.y = __uint_as_float(r[4 * i + 1]) + consts[c + 1] + res[i].y;
            v.z = __uint_as_float(r[4 * i + 2]) + consts[c + 2] + res[i].z;
            v.w = __uint_as_float(r[4 * i + 3]) + consts[c + 3] + res[i].w;
            myrow[c >> 2] = v;
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
          }
        }
      };
      auto residual = [&](float4 (&e)[4], int c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = c0 + 4 * i;
          e[i] = c < d ? myrow[c >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // two 16-column units per TMEM round trip; the residual loads are issued before the round trip completes
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 32) {
        uint32_t ra[16], rb[16];
        float4 e[4];
        constexpr bool kTwoAlways = (DP % 32) == 0;
        const bool two = kTwoAlways || c0 + 16 < DP;
        FFD_TMEM_LD16(tmem_base + lane_base + COL_Y + s * 128 + c0, ra);
        if (two) FFD_TMEM_LD16(tmem_base + lane_base + COL_Y + s * 128 + c0 + 16, rb);
        residual(e, c0);
        tmem_ld_wait();
        if (c0 + 32 >= DP) {
          tc_fence_before();
          mbar_arrive_cta0(&y_empty[s]);
        }
        unit(ra, e, c0);
        if (two) {
          residual(e, c0 + 16);
          unit(rb, e, c0 + 16);
        }
      }
      const float mean = s1 / (float)d;
      const float rstd = rsqrtf(fmaxf(s2 / (float)d - mean * mean, 0.f) + eps);
      __syncwarp();  // this warp's 32 rows are in the buffer
      {
        // sweep 2: the warp's 32 rows are one contiguous block of the tile; lane l takes its 16-byte chunks l,
        // l + 32, ... (coalesced global stores); the row statistics come from the lane that owns the row
        const int wrow0 = quad * 32;
        const float4* wv = vb + (size_t)wrow0 * n4;
        float4* dst = reinterpret_cast<float4*>(y + (row0 + wrow0) * d);
        const int wrows = rows - wrow0 < 32 ? rows - wrow0 : 32;  // may be <= 0 in the ragged last tile
        const int step_r = 32 / n4, step_c = 32 % n4;
        int r = lane / n4, c4 = lane % n4;
        for (int q = lane; q < 32 * n4; q += 32) {
          const float m_r = __shfl_sync(0xffffffffu, mean, r & 31);
          const float s_r = __shfl_sync(0xffffffffu, rstd, r & 31);
          if (r < wrows) {
            const float4 v = wv[q];
            const float4 w4 = *reinterpret_cast<const float4*>(consts + DP + 4 * c4);
            const float4 b4 = *reinterpret_cast<const float4*>(consts + 2 * DP + 4 * c4);
            float4 o;
            o.x = (v.x - m_r) * s_r * w4.x + b4.x;
            o.y = (v.y - m_r) * s_r * w4.y + b4.y;
            o.z = (v.z - m_r) * s_r * w4.z + b4.z;
            o.w = (v.w - m_r) * s_r * w4.w + b4.w;
            if (eps < 0.f) o = v;  // debug: pre-LayerNorm values (FFD_FFN_RAW)
            dst[q] = o;
          }
          r += step_r, c4 += step_c;
          if (c4 >= n4) c4 -= n4, ++r;
        }
      }
      fence_proxy_async();  // the next bulk copy into this buffer is ordered after these reads
      mbar_arrive(&y_ready[(np & 1) * 2 + s]);
    };
    stage_x(0);
    for (uint32_t np = 0; np < my_pairs; ++np) {
      FFN_STAMP(warp == 4 && np == 2, 200);
      if (np + 1 < my_pairs) stage_x(np + 1);
      FFN_STAMP(warp == 4 && np == 2, 201);
      finish(np);
      FFN_STAMP(warp == 4 && np == 2, 202);
    }
  } else {
    // ================= epilogue groups (slices only): tile s = g & 1, column half hf = g >> 1 =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    const int g = (warp - 12) >> 2;
    const int s = g & 1, hf = g >> 1;
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t hcol = tmem_base + lane_base + COL_H + s * 128 + hf * 64;  // this group's 64 fp32 columns
    const uint32_t n_slices = my_pairs * (uint32_t)n_chunks;
    // this group's half of H_s (TMEM fp32) -> relu -> bf16, written back in place
    for (uint32_t nchunk = 0; nchunk < n_slices; ++nchunk) {
      mbar_wait_relaxed(&h_full[s], nchunk & 1, 20);
      tc_fence_after();
      FFN_STAMP(warp == 12 && nchunk - 2u * n_chunks < (uint32_t)n_chunks, (nchunk - 2u * n_chunks) * 8 + 3);
      uint32_t r[64], pk[32];
      FFD_TMEM_LD16(hcol, r);
      FFD_TMEM_LD16(hcol + 16, (r + 16));
      FFD_TMEM_LD16(hcol + 32, (r + 32));
      FFD_TMEM_LD16(hcol + 48, (r + 48));
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) pk[e] = relu_pack_op(r[2 * e], r[2 * e + 1]);
      // 64 fp32 columns -> 32 packed columns at the start of this group's own half
      FFD_TMEM_ST16(hcol, pk);
      FFD_TMEM_ST16(hcol + 16, (pk + 16));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_cta0(&p_full[s]);
      FFN_STAMP(warp == 12 && nchunk - 2u * n_chunks < (uint32_t)n_chunks, (nchunk - 2u * n_chunks) * 8 + 4);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves while the pair's MMAs may still read its shared memory / TMEM
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

// ---- weight packing ----------------------------------------------------------------------------------
// Per layer: n_chunks W1 images [128 n x DP k] then n_chunks W2 images [DP n x 128 k], bf16, K-major
// core-matrix order: byte(n, k) = (n/8)*128 + (k/8)*(rows*16) + (n%8)*16 + (k%8)*2.
__global__ void pack_ffn_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                const float* __restrict__ W2, op16* __restrict__ out, int d, int F, int DP) {
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
        v = round_op(b1[f]);
      } else if (k == d + 1) {
        v = b1[f] - round_op(b1[f]);
      }
    } else {
      const int f = j * CHUNK + k;  // n = output feature, k = hidden unit inside the slice
      if (n < d) v = W2[(int64_t)n * F + f];
    }
    out[i] = to_op(v);
  }
}

// Per-CTA halves of the weight slices for the CTA-pair kernel, from the full images: per slice j two W1 halves
// [64 n x DP k] (rows 64 r .. 64 r + 63 of the slice) and two W2 halves [DP/2 n x 128 k] (output features DP/2 r ..),
// each a self-contained K-major core-matrix image.
__global__ void pack_ffn_halves_kernel(const op16* __restrict__ full, op16* __restrict__ out, int F,
                                       int DP) {
  const int n_chunks = F / CHUNK;
  const int64_t per_img = (int64_t)CHUNK * DP;
  const int64_t total = 2 * n_chunks * per_img;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const bool is_w2 = i >= n_chunks * per_img;
    const int64_t r = is_w2 ? i - n_chunks * per_img : i;
    const int j = (int)(r / per_img);
    const int e = (int)(r % per_img);              // element inside [half 0 | half 1]
    const int hrows = is_w2 ? DP / 2 : CHUNK / 2;  // rows of a half image
    const int half = e / (int)(per_img / 2), eh = e % (int)(per_img / 2);
    const int kg = eh / (hrows * 8), rem = eh % (hrows * 8);
    const int n = half * hrows + (rem / 64) * 8 + (rem % 64) / 8, k = kg * 8 + rem % 8;
    const int rows = is_w2 ? DP : CHUNK;           // rows of the full image
    const int64_t src = (is_w2 ? n_chunks * per_img : 0) + (int64_t)j * per_img + (int64_t)kg * rows * 8 + (n / 8) * 64 +
                        (n % 8) * 8 + k % 8;
    out[i] = full[src];
  }
}

}  // namespace tc

bool tc_supported(const ffd_dims& m) {
  return m.d_model + 2 <= 80 && m.d_model % 4 == 0 && m.d_ff % tc::CHUNK == 0 && m.d_ff >= tc::CHUNK;
}

int64_t tc_ffn_image_bytes(const ffd_dims& m) {
  const int DP = tc::dp_of(m.d_model);
  return (int64_t)2 * (m.d_ff / tc::CHUNK) * tc::CHUNK * DP * 2;
}

// per layer: [FFN W1 images][FFN W2 images][WqP image][WoP image]
// per layer: [FFN W1 / W2 images][WqP WoP (WkP WvP)][FFN half images of the CTA-pair kernel]
static int64_t halves_offset(const ffd_dims& m) {
  return tc_ffn_image_bytes(m) + (tc_attn_supported(m) ? tc_attn_image_bytes(m) : 0);
}
static int64_t layer_image_bytes(const ffd_dims& m) { return halves_offset(m) + tc_ffn_image_bytes(m); }

int64_t tc_packed_bytes(const ffd_dims& m) {
  if (!tc_supported(m)) return 0;
  return layer_image_bytes(m) * m.n_layers;
}

const void* tc_ffn_image(const ffd_dims& m, const void* packed, int layer) {
  return (const uint8_t*)packed + (size_t)layer * layer_image_bytes(m);
}

const void* tc_ffn_half_image(const ffd_dims& m, const void* packed, int layer) {
  return (const uint8_t*)packed + (size_t)layer * layer_image_bytes(m) + halves_offset(m);
}

const void* tc_attn_image(const ffd_dims& m, const void* packed, int layer) {
  return (const uint8_t*)packed + (size_t)layer * layer_image_bytes(m) + tc_ffn_image_bytes(m);
}

int tc_pack_weights(const ffd_dims& m, const float* weights, void* packed, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const int DP = tc::dp_of(m.d_model);
  for (int l = 0; l < m.n_layers; ++l) {
    const float* w = weights + ar.layer0 + (int64_t)l * ar.layer_stride;
    tc::pack_ffn_kernel<<<256, 256, 0, s>>>(w + ar.w1, w + ar.b1, w + ar.w2,
                                            (tc::op16*)((uint8_t*)packed + l * layer_image_bytes(m)),
                                            m.d_model, m.d_ff, DP);
    FFD_CHECK_LAUNCH("pack_ffn_kernel");
    tc::op16* full = (tc::op16*)((uint8_t*)packed + l * layer_image_bytes(m));
    tc::pack_ffn_halves_kernel<<<256, 256, 0, s>>>(full, (tc::op16*)((uint8_t*)full + halves_offset(m)), m.d_ff, DP);
    FFD_CHECK_LAUNCH("pack_ffn_halves_kernel");
  }
  if (tc_attn_supported(m)) return tc_pack_attn(m, weights, packed, layer_image_bytes(m), tc_ffn_image_bytes(m), s);
  return FFD_OK;
}

extern long long* g_attn_dbg;  // optional clock64 stamp buffer (ffd_debug_set_attn_stamps)

// LayerNorm eps; FFD_FFN_RAW=1 (debug) makes both kernels store the pre-LayerNorm values instead
static float ffn_eps() {
  static const float eps = getenv("FFD_FFN_RAW") ? -1.f : 1e-5f;
  return eps;
}

template <int DP>
static int launch_ffn_tc(const ffd_dims& m, const uint8_t* img, const float* w, const Arena& ar, const float* x,
                         float* y, int64_t M, cudaStream_t s) {
  const tc::FfnSmem sm = tc::ffn_smem(DP, m.d_model);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    FFD_CHECK_CUDA(cudaGetDevice(&dev));
    FFD_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::ffn_ln_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      sm.total));
  const int n_chunks = m.d_ff / tc::CHUNK;
  const int64_t n_pairs = (M + 2 * tc::TILE_M - 1) / (2 * tc::TILE_M);
  const unsigned grid = (unsigned)(n_pairs < n_sm ? n_pairs : n_sm);
  const uint8_t* w1img = img;
  const uint8_t* w2img = img + (size_t)n_chunks * sm.w1_bytes;
  tc::ffn_ln_tc_kernel<DP><<<grid, tc::NUM_THREADS, sm.total, s>>>(x, y, w1img, w2img, w + ar.b2, w + ar.n2w,
                                                                    w + ar.n2b, M, m.d_model, n_chunks, n_pairs,
                                                                    ffn_eps(), g_attn_dbg);
  FFD_CHECK_LAUNCH("ffn_ln_tc_kernel");
  return FFD_OK;
}

// The CTA-pair (cta_group::2) kernel: two SMs share every weight slice.  FFD_FFN_CTA2=0 / 1 forces the choice.
template <int DP, int D>
static int launch_ffn_tc2(const ffd_dims& m, const uint8_t* himg, const float* w, const Arena& ar, const float* x,
                          float* y, int64_t M, cudaStream_t s) {
  const tc::FfnSmem sm = tc::ffn2_smem(DP, m.d_model);
  static int n_clusters = 0;
  const int n_chunks = m.d_ff / tc::CHUNK;
  if (n_clusters == 0) {
    FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::ffn_ln_tc2_kernel<DP, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        sm.total));
    int dev = 0, n_sm = 0;
    FFD_CHECK_CUDA(cudaGetDevice(&dev));
    FFD_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_sm / 2 * 2));
    cfg.blockDim = dim3(tc::NUM_THREADS);
    cfg.dynamicSmemBytes = sm.total;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
    cfg.attrs = &at, cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, tc::ffn_ln_tc2_kernel<DP, D>, &cfg) != cudaSuccess || nc <= 0) {
      (void)cudaGetLastError();
      nc = n_sm / 2;
    }
    n_clusters = nc < n_sm / 2 ? nc : n_sm / 2;
    if (getenv("FFD_DEBUG")) fprintf(stderr, "ffn pairs: max active clusters %d -> %d clusters\n", nc, n_clusters);
  }
  const int64_t n_pairs = (M + 2 * tc::TILE_M - 1) / (2 * tc::TILE_M);
  const int64_t n_quads = (n_pairs + 1) / 2;
  const unsigned grid = 2u * (unsigned)(n_quads < n_clusters ? n_quads : n_clusters);
  const uint8_t* w1h = himg;
  const uint8_t* w2h = himg + (size_t)n_chunks * 2 * sm.w1_bytes;
  tc::ffn_ln_tc2_kernel<DP, D><<<grid, tc::NUM_THREADS, sm.total, s>>>(x, y, w1h, w2h, w + ar.b2, w + ar.n2w, w + ar.n2b,
                                                                     M, m.d_model, n_chunks, n_pairs, ffn_eps(),
                                                                     g_attn_dbg);
  FFD_CHECK_LAUNCH("ffn_ln_tc2_kernel");
  return FFD_OK;
}

static int g_ffn_pairs_mode = -1;  // -1 unknown (FFD_FFN_CTA2 or the default), 0 off, 1 on
void tc_debug_set_ffn_pairs(int mode) { g_ffn_pairs_mode = mode; }

static bool ffn_use_pairs(const ffd_dims& m) {
  if (g_ffn_pairs_mode < 0) {
    const char* e = getenv("FFD_FFN_CTA2");
    g_ffn_pairs_mode = e ? (e[0] == '1') : FFD_FFN_CTA2_DEFAULT;
  }
  return g_ffn_pairs_mode == 1 && m.d_model == 72;  // the default model (the pair kernel is instantiated for it)
}

int tc_ffn_ln(const ffd_dims& m, const void* packed, const float* weights, int layer, const float* x, float* y,
              int64_t M, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const uint8_t* img = (const uint8_t*)packed + (size_t)layer * layer_image_bytes(m);
  if (ffn_use_pairs(m)) return launch_ffn_tc2<80, 72>(m, img + halves_offset(m), w, ar, x, y, M, s);
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
