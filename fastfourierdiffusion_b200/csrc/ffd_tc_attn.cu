// tcgen05 / TMEM kernel for the attention half of a PURE-CACHE encoder layer (998 of 1000 default steps):
//   y = LN1(x + out_proj(softmax(q K_c^T / sqrt(hd)) V_c)),  q = x Wq^T + bq
//   cached_transformer.py:228-258 (pure-cache K/V), :309-321 (attention, out-proj, residual, norm1)
// K_c / V_c are the sample-0 cache rows shared by the whole batch (caching.py:326-328), so every token
// row is independent and the work is a chain of GEMMs over 128-token tiles:
//   Q[128 x QN]    = X[128 x DP] . WqP^T          G = 15 / hd heads share one 16-column tile (head g of a tile
//                                                 sits in slots [g hd, (g+1) hd), slot 15 is the constant 1);
//                                                 log2(e)/sqrt(hd) and bq are folded in
//   S_v[128 x CK]  = Qtile(h) . K_v^T             one MMA per (head h, key chunk c) = "virtual head" v; the K
//                                                 image of head h is zero in the slots of its tile mates
//   P_v            = exp2(S_v - m)                softmax warps: TMEM -> registers -> TMEM (fp16 pairs).
//                                                 The CK key columns are split in four quarters, one per softmax
//                                                 group; each quarter uses ITS OWN row maximum as the shift and
//                                                 gets its own O accumulator; quarters (and key chunks) are
//                                                 combined when O is read out (flash-style split-K: exact)
//   O_v[128 x 16]  = P_v . [V_v | 1]              A operand read from TMEM; the ones row of the V image makes
//                                                 the tensor core return the softmax row sum as well
//   OUT[128 x DP]  = concat_h(O_h / rowsum_h) . WoP^T   (bo folded in) -> + residual -> LayerNorm
// The fp32 token tile enters and leaves shared memory as one bulk async copy each way; the [tokens x heads x L]
// score tensor the reference materialises twice per layer (27.5 GB each at B=16384) lives in TMEM.
//
// v11 pipeline.  The softmax warps do nothing but heads, and nothing they wait for is issued late:
//  * ACROSS tiles: while they work on tile t, four "tile warps" finish tile t-1 (out-proj read-out, residual,
//    LayerNorm) and prepare tile t+1 (bf16 staging of X, Q projection read-out) through a 64-column TMEM scratch
//    region, so the MUFU pipe does not drain at a tile boundary.
//  * INSIDE a head the next head's scores are loaded and reduced to their row maximum while the current head's
//    exponentials are still in the MUFU queue.  The scores live in ONE buffer that is free again as soon as every
//    softmax warp holds its share in registers; P goes to its own double buffer.  The score MMA of head n+2 and
//    the P.V MMA of head n are issued by two different warps, so neither waits behind the other's barrier.
//
// Warp roles (28 warps, register budgets per warpgroup via setmaxnreg): 0 = K/V image + weight producer,
// 1 = TMEM allocator + score MMAs, 2 = token-tile load/store, 3 = P.V MMAs, 4..7 = O read-out (combine
// quarters/chunks, normalise, write the out-proj operand), 8..11 = tile warps (warp 8 also issues their scratch
// MMAs: Q projection / out-projection), 12..27 = four softmax groups of four warps (88 registers: they hold two
// heads' scores).  One eighth of the exponentials runs on the FMA pipe (FFD_POLY_PAIRS).
//
// Lown > 0 runs STOCK attention with the same code: tiles are aligned to samples and every sample brings its own
// K/V images (kv_own_tc_kernel below builds them with the K/V projection fused in).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ffd_tc_ptx.cuh"

namespace ffd {
namespace tc {

constexpr int ATT_THREADS = 896;
constexpr int ATT2_THREADS = 1024;  // v12: + one warpgroup of four MMA-issuing warps
constexpr int K_STAGES = 3;    // K image ring
constexpr int V_STAGES = 2;    // V image ring
constexpr int HP = 16;         // slots per Q tile (UMMA K of the score MMA, N of the PV MMA)
constexpr int SCR_COLS = 64;   // TMEM scratch columns (Q projection / out-proj passes)
// Share of the softmax exponentials evaluated on the FMA pipe instead of MUFU: bit mask over the 8 fp16 pairs of
// a 16-column chunk.  Same-box sweep (tools/ab.sh, ECG shapes): 0x00 2.031 ms, 0x01 (1/8) 1.980, 0x81 / 0x03 (2/8)
// 2.013, 0x11 2.021, 0x49 (3/8) 2.075, 0x55 (4/8) 2.176 -- the loop is bound by issue slots as much as by MUFU, so
// only a small share pays.
#ifndef FFD_POLY_PAIRS
#define FFD_POLY_PAIRS 0x01u
#endif
#ifndef FFD_STAGGER_NS
#define FFD_STAGGER_NS 250   // start-up offset between the four softmax warps of a sub-partition
#endif
#ifndef FFD_ATTN_RING
#define FFD_ATTN_RING 0  // v12: 1 / 2 = pass the exponential phase round the four softmax groups / two sets of two
#endif
constexpr uint32_t POLY_PAIRS = FFD_POLY_PAIRS;

// TMEM columns
constexpr uint32_t COL_S = 0;       // scores of the current virtual head (up to 192 fp32 columns)
constexpr uint32_t COL_P0 = 192;    // P (fp16 pairs, up to 96 columns) of the even virtual heads
constexpr uint32_t COL_P1 = 288;    // ... of the odd ones
constexpr uint32_t COL_O = 384;     // O accumulators: quarter q at COL_O + 16 q
constexpr uint32_t COL_SCR = 448;   // scratch: Q projection / out-proj accumulators, 64 columns per pass

struct AttnSmem {
  int x32, xo, obf, qbf, wq, wo, kring, vring, consts, mq, bars, tmem, total;
  int tile_bytes, tile_stride, k_bytes, v_bytes, wq_bytes, wo_bytes, qbf_bytes;
};

__host__ __device__ inline AttnSmem attn_smem(int DP, int QN, int CK, int d) {
  AttnSmem s;
  int o = 0;
  auto take = [&](int bytes) {
    int r = o;
    o += (bytes + 127) / 128 * 128;
    return r;
  };
  s.tile_bytes = TILE_M * d * 4;
  s.tile_stride = (s.tile_bytes + 127) / 128 * 128;
  s.k_bytes = CK * HP * 2;
  s.v_bytes = HP * CK * 2;
  s.wq_bytes = QN * DP * 2;
  s.wo_bytes = DP * DP * 2;
  s.qbf_bytes = TILE_M * QN * 2;
  s.x32 = take(2 * s.tile_stride);      // fp32 token tiles (two in flight): input, residual, output staging
  s.xo = take(TILE_M * DP * 2);         // bf16 A operand of the Q projection (X of the next tile)
  s.obf = take(TILE_M * DP * 2);        // bf16 A operand of the out-projection (normalised O of this tile)
  s.qbf = take(2 * s.qbf_bytes);        // bf16 Q tiles of this tile and the next
  s.wq = take(s.wq_bytes);
  s.wo = take(s.wo_bytes);
  s.kring = take(K_STAGES * s.k_bytes);
  s.vring = take(V_STAGES * s.v_bytes);
  s.consts = take(2 * DP * 4);
  s.mq = take(4 * 4 * TILE_M * 4);      // quarter row maxima [virtual head & 3][quarter][row]
  s.bars = take(40 * 8);
  s.tmem = take(16);
  s.total = o;
  return s;
}

template <int DP, int KQ>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_cached_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ wqimg,
                      const uint8_t* __restrict__ woimg, const uint8_t* __restrict__ kvimg,
                      const float* __restrict__ lnw, const float* __restrict__ lnb, int64_t M, int d, int H, int hd,
                      int G, int QN, int NCH, int64_t n_tiles, float eps, int Lown, int tps,
                      long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int CK = 64 * KQ;   // keys per chunk
  constexpr int NQ = 16 * KQ;   // key columns per softmax group
  constexpr int NP = 8 * KQ;    // packed P columns per softmax group
  const int NV = H * NCH;       // virtual heads per tile
  const AttnSmem sm = attn_smem(DP, QN, CK, d);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* k_full = bars;                       // [K_STAGES] tx
  uint64_t* k_empty = k_full + K_STAGES;         // [K_STAGES] commit
  uint64_t* v_full = k_empty + K_STAGES;         // [V_STAGES] tx
  uint64_t* v_empty = v_full + V_STAGES;         // [V_STAGES] commit
  uint64_t* w_full = v_empty + V_STAGES;         // [1] tx (Wq + Wo, once)
  uint64_t* t_full = w_full + 1;                 // [2] tx: fp32 tile landed
  uint64_t* y_ready = t_full + 2;                // [2] 128 arrivals: LayerNorm output written to the tile buffer
  uint64_t* scr_full = y_ready + 2;              // [1] commit: a scratch pass has been computed
  uint64_t* qbf_full = scr_full + 1;             // [2] 128 arrivals: bf16 Q tiles of a token tile written
  uint64_t* s_full = qbf_full + 2;               // [1] commit: scores of a virtual head computed
  uint64_t* s_empty = s_full + 1;                // [1] 512 arrivals: every softmax warp holds them in registers
  uint64_t* p_full = s_empty + 1;                // [2] 512 arrivals (all four groups)
  uint64_t* p_empty = p_full + 2;                // [2] commit: the P.V MMA has read the buffer
  uint64_t* o_full = p_empty + 2;                // [1] commit
  uint64_t* o_empty = o_full + 1;                // [1] 128 arrivals (read-out warps)
  uint64_t* obf_full = o_empty + 1;              // [1] 128 arrivals (read-out warps)
  uint64_t* obf_empty = obf_full + 1;            // [1] commit: the out-projection has read its operand
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.tmem);
  float* consts = reinterpret_cast<float*>(smem + sm.consts);
  float* mq = reinterpret_cast<float*>(smem + sm.mq);   // [4][4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t my_tiles = (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // >= 1
  // Rows of this CTA's t-th tile.  Lown == 0: K/V are the shared cache, tiles are consecutive 128-row blocks of
  // [0, M).  Lown > 0 (stock attention, cached_transformer.py:123-134 / nn.MultiheadAttention): every sample
  // attends to its OWN keys, so tiles are aligned to samples (tps tiles of <= 128 rows per sample of Lown rows)
  // and the K/V images are per sample.
  auto tile_rows = [&](uint32_t t, int64_t& row0, int& rows, int64_t& sample) {
    const int64_t gt = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
    if (Lown == 0) {
      sample = 0;
      row0 = gt * TILE_M;
      rows = (int)(M - row0 < TILE_M ? M - row0 : TILE_M);
    } else {
      sample = gt / tps;
      const int part = (int)(gt - sample * tps);
      row0 = sample * Lown + (int64_t)part * TILE_M;
      rows = Lown - part * TILE_M < TILE_M ? Lown - part * TILE_M : TILE_M;
    }
  };
  const uint32_t n_heads = my_tiles * (uint32_t)NV;  // virtual heads this CTA runs through
#ifdef FFD_ATTN_STAMPS  // timeline of local tile 2 of CTA 0 (tools/attn_stamps.py)
#define FFD_STAMP(cond, slot) do { if (dbg && blockIdx.x == 0 && lane == 0 && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define FFD_STAMP(cond, slot) do { } while (0)
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < K_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < V_STAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&t_full[b], 1);
      mbar_init(&y_ready[b], 128);
      mbar_init(&qbf_full[b], 128);
      mbar_init(&p_full[b], 512);
      mbar_init(&p_empty[b], 1);
    }
    mbar_init(scr_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 512);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 128);
    mbar_init(obf_full, 128);
    mbar_init(obf_empty, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * DP; i += ATT_THREADS) {
    const int w = i / DP, c = i % DP;
    consts[i] = c < d ? (w == 0 ? lnw[c] : lnb[c]) : 0.f;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // provably warp-uniform

  // register budget per warpgroup (896 threads x 72 at launch): the softmax warps hold two heads' scores
  // (set at the top of each role branch so that ptxas allocates per branch)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ================= K/V image producer =================
    if (lane == 0) {
      mbar_expect_tx(w_full, sm.wq_bytes + sm.wo_bytes);
      bulk_load(smem + sm.wq, wqimg, sm.wq_bytes, w_full);
      bulk_load(smem + sm.wo, woimg, sm.wo_bytes, w_full);
      uint32_t nk = 0, nv = 0;
      const size_t img_stride = (size_t)sm.k_bytes + sm.v_bytes;
      const size_t sample_stride = Lown ? (size_t)NV * img_stride : 0;
      auto load_k = [&](const uint8_t* base, uint32_t v) {
        const uint32_t st = nk % K_STAGES, u = nk / K_STAGES;
        mbar_wait_relaxed(&k_empty[st], (u & 1) ^ 1, 200);
        mbar_expect_tx(&k_full[st], sm.k_bytes);
        bulk_load(smem + sm.kring + st * sm.k_bytes, base + v * img_stride, sm.k_bytes, &k_full[st]);
        ++nk;
      };
      auto load_v = [&](const uint8_t* base, uint32_t v) {
        const uint32_t st = nv % V_STAGES, u = nv / V_STAGES;
        mbar_wait_relaxed(&v_empty[st], (u & 1) ^ 1, 200);
        mbar_expect_tx(&v_full[st], sm.v_bytes);
        bulk_load(smem + sm.vring + st * sm.v_bytes, base + v * img_stride + sm.k_bytes, sm.v_bytes, &v_full[st]);
        ++nv;
      };
      auto tile_base = [&](uint32_t t) {
        int64_t row0, sample;
        int rows;
        tile_rows(t, row0, rows, sample);
        return kvimg + (size_t)sample * sample_stride;
      };
      // K images run two virtual heads ahead of the V images (the score MMAs run ahead of the P.V MMAs)
      uint32_t vk = 0, vv = 0, tk = 0, tv = 0;  // virtual head inside its tile / tile of the next K and V image
      const uint8_t* kbase = tile_base(0);
      const uint8_t* vbase = kbase;
      auto next_k = [&]() {
        load_k(kbase, vk);
        if (++vk == (uint32_t)NV) {
          vk = 0;
          if (++tk < my_tiles) kbase = tile_base(tk);
        }
      };
      next_k();
      if (n_heads > 1) next_k();
      for (uint32_t n = 0; n < n_heads; ++n) {
        load_v(vbase, vv);
        if (++vv == (uint32_t)NV) {
          vv = 0;
          if (++tv < my_tiles) vbase = tile_base(tv);
        }
        if (n + 2 < n_heads) next_k();
      }
    }
  } else if (warp == 2) {
    // ================= token-tile load / store =================
    if (lane == 0) {
      auto load_tile = [&](uint32_t t) {
        const uint32_t b = t & 1;
        int64_t row0, sample;
        int rows;
        tile_rows(t, row0, rows, sample);
        mbar_expect_tx(&t_full[b], (uint32_t)(rows * d * 4));
        bulk_load(smem + sm.x32 + b * sm.tile_stride, x + row0 * d, (uint32_t)(rows * d * 4), &t_full[b]);
      };
      load_tile(0);
      if (my_tiles > 1) load_tile(1);
      for (uint32_t t = 0; t < my_tiles; ++t) {
        const uint32_t b = t & 1;
        int64_t row0, sample;
        int rows;
        tile_rows(t, row0, rows, sample);
        mbar_wait_relaxed(&y_ready[b], (t >> 1) & 1, 100);
        bulk_store(y + row0 * d, smem + sm.x32 + b * sm.tile_stride, (uint32_t)(rows * d * 4));
        bulk_commit();
        bulk_wait_read();
        if (t + 2 < my_tiles) load_tile(t + 2);  // the buffer is free again
      }
      bulk_wait_all();
    }
  } else if (warp == 1) {
    // ================= score MMAs (whole warp converged, one elected lane issues) ===========
    const uint32_t idesc_s = make_idesc(TILE_M, CK);
    constexpr uint32_t LBO128 = 128 * 16;
    const uint32_t qbf = smem_u32(smem + sm.qbf);
    const uint32_t kring = smem_u32(smem + sm.kring);
    uint32_t t = 0, nk = 0;
    int v = 0, c = 0, g = 0, qt = 0;  // position of the score MMA being issued: tile, virtual head, chunk, place
    for (uint32_t m = 0; m < n_heads; ++m) {
      if (m > 0) mbar_wait_relaxed(s_empty, (m - 1) & 1, 40);  // all softmax warps hold S_{m-1} in registers
      if (v == 0) mbar_wait_relaxed(&qbf_full[t & 1], (t >> 1) & 1, 60);  // this tile's Q tiles are in shared memory
      const uint32_t st = nk % K_STAGES;
      mbar_wait(&k_full[st], (nk / K_STAGES) & 1);
      tc_fence_after();
      umma(tmem_base + COL_S, make_desc(qbf + (t & 1) * sm.qbf_bytes + qt * 2 * LBO128, LBO128, 128),
           make_desc(kring + st * sm.k_bytes, CK * 16, 128), idesc_s, 0);
      umma_commit(s_full);
      umma_commit(&k_empty[st]);
      ++nk;
      FFD_STAMP(t == 2, v * 2 + 0);
      if (++c == NCH) {
        c = 0;
        if (++g == G) {
          g = 0;
          ++qt;
        }
      }
      if (++v == NV) {
        v = 0, c = 0, g = 0, qt = 0;
        ++t;
      }
    }
    mbar_wait_relaxed(s_full, (n_heads - 1) & 1, 200);  // drain: nothing of this warp is in flight when the CTA exits
  } else {
    // ================= P.V MMAs (warp 3) =================
    constexpr uint32_t IDESC_PV = make_idesc_f16(TILE_M, HP);
    const uint32_t vring = smem_u32(smem + sm.vring);
    for (uint32_t n = 0; n < n_heads; ++n) {
      const uint32_t b = n & 1;
      const uint32_t st = n % V_STAGES;
      mbar_wait_relaxed(&p_full[b], (n >> 1) & 1, 32);   // P_n (fp16 pairs) is complete
      FFD_STAMP(n - 2u * NV < (uint32_t)NV, 64 + (n - 2u * NV) * 2 + 0);
      mbar_wait_relaxed(o_empty, (n & 1) ^ 1, 20);       // O of the previous virtual head has been read out
      mbar_wait(&v_full[st], (n / V_STAGES) & 1);
      tc_fence_after();
      const uint32_t pcol = tmem_base + (b ? COL_P1 : COL_P0);
      // one chain of KQ K-steps per quarter of the keys, each quarter into its own accumulator
      const uint64_t vdesc = make_desc(vring + st * sm.v_bytes, HP * 16, 128);
      if constexpr (KQ == 3) {
        umma_ts_4x3(tmem_base + COL_O, pcol, NP, vdesc, (2 * HP * 16) >> 4, IDESC_PV);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          umma_ts_chain<KQ>(tmem_base + COL_O + 16 * q, pcol + q * NP, 8,
                            vdesc + (uint64_t)(q * KQ) * ((2 * HP * 16) >> 4), (2 * HP * 16) >> 4, IDESC_PV, 0);
      }
      umma_commit(o_full);
      umma_commit(&p_empty[b]);
      umma_commit(&v_empty[st]);
      FFD_STAMP(n - 2u * NV < (uint32_t)NV, 64 + (n - 2u * NV) * 2 + 1);
    }
    mbar_wait_relaxed(&p_empty[(n_heads - 1) & 1], ((n_heads - 1) >> 1) & 1, 200);  // drain
  }
  } else if (warp < 8) {
    // ================= O read-out warps =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    uint8_t* ob = smem + sm.obf;
    // constant columns of the out-proj operand: d, d+1 = 1 (bias hi/lo), the rest of the padding = 0
    for (int c = d; c < DP; ++c)
      *reinterpret_cast<op16*>(ob + row_off + (c >> 3) * (TILE_M * 16) + (c & 7) * 2) =
          to_op((c == d || c == d + 1) ? 1.f : 0.f);
    uint32_t n = 0;
    for (uint32_t t = 0; t < my_tiles; ++t) {
      for (int h = 0; h < H; ++h) {
        // With hd <= 7 the row sum sits in slot 7 (pack_kv_kernel), so the low 8 columns of every quarter
        // accumulator hold everything: ONE TMEM round trip per virtual head.  (The P.V MMA of the next head waits
        // for this read: the round trip is on the kernel's critical loop.)
        if (hd <= 7) {
          float acc[8];
          float mrun = 0.f;
          for (int c = 0; c < NCH; ++c, ++n) {
            mbar_wait_relaxed(o_full, n & 1, 32);
            tc_fence_after();
            FFD_STAMP(warp == 4 && t == 2, 256 + (h * NCH + c) * 2 + 0);
            uint32_t r[32];
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O, r);
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O + 16, (r + 8));
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O + 32, (r + 16));
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O + 48, (r + 24));
            // combine the four quarters: weights exp2(m_q - max m_q)
            const float* mh = mq + ((n & 3) * 4) * TILE_M + row;
            const float q0 = mh[0], q1 = mh[TILE_M], q2 = mh[2 * TILE_M], q3 = mh[3 * TILE_M];
            float mm = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
            if (c > 0) mm = fmaxf(mm, mrun);
            const float w0 = exp2_poly3(q0 - mm), w1 = exp2_poly3(q1 - mm), w2 = exp2_poly3(q2 - mm),
                        w3 = exp2_poly3(q3 - mm);
            const float wp = c > 0 ? exp2_poly3(mrun - mm) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = c > 0 ? acc[j] * wp : 0.f;
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty);
            FFD_STAMP(warp == 4 && t == 2, 256 + (h * NCH + c) * 2 + 1);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w0, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[8 + j]), __uint_as_float(r[9 + j]), w1, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[16 + j]), __uint_as_float(r[17 + j]), w2, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[24 + j]), __uint_as_float(r[25 + j]), w3, acc[j], acc[j + 1]);
            }
            mrun = mm;
          }
          const float inv = rcp_fma(acc[7]);
          // the out-projection of the previous tile must have read its operand before the first overwrite
          if (h == 0 && t >= 1) mbar_wait_relaxed(obf_empty, (t - 1) & 1, 60);
          if ((hd & 1) == 0) {  // even head dim: the head's columns start even, two columns per store
#pragma unroll
            for (int j = 0; j < 6; j += 2) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<uint32_t*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    pack_op(acc[j] * inv, acc[j + 1] * inv);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<op16*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    to_op(acc[j] * inv);
              }
            }
          }
        } else {
          float acc[HP];
          float mrun = 0.f;
          for (int c = 0; c < NCH; ++c, ++n) {
            mbar_wait_relaxed(o_full, n & 1, 32);
            tc_fence_after();
            uint32_t r[32];
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O, r);
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 16, (r + 16));
            const float* mh = mq + ((n & 3) * 4) * TILE_M + row;
            const float q0 = mh[0], q1 = mh[TILE_M], q2 = mh[2 * TILE_M], q3 = mh[3 * TILE_M];
            float mm = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
            if (c > 0) mm = fmaxf(mm, mrun);
            const float w0 = exp2_poly3(q0 - mm), w1 = exp2_poly3(q1 - mm), w2 = exp2_poly3(q2 - mm),
                        w3 = exp2_poly3(q3 - mm);
            const float wp = c > 0 ? exp2_poly3(mrun - mm) : 0.f;
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[j] = c > 0 ? acc[j] * wp : 0.f;
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
              fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w0, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[16 + j]), __uint_as_float(r[17 + j]), w1, acc[j], acc[j + 1]);
            }
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 32, r);
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 48, (r + 16));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty);
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
              fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w2, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[16 + j]), __uint_as_float(r[17 + j]), w3, acc[j], acc[j + 1]);
            }
            mrun = mm;
          }
          const float inv = rcp_fma(acc[HP - 1]);
          if (h == 0 && t >= 1) mbar_wait_relaxed(obf_empty, (t - 1) & 1, 60);
          if ((hd & 1) == 0) {
#pragma unroll
            for (int j = 0; j < HP - 2; j += 2) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<uint32_t*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    pack_op(acc[j] * inv, acc[j + 1] * inv);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < HP - 1; ++j) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<op16*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    to_op(acc[j] * inv);
              }
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(obf_full);
    }
  } else if (warp < 12) {
    // ================= tile warps: prepare tile t+1 / finish tile t-1 while the heads of tile t run ==========
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    const uint32_t scr = tmem_base + lane_base + COL_SCR;
    uint8_t* xo = smem + sm.xo;
    const bool lead = warp == 8;  // issues the scratch MMAs of the group (whole warp converged)
    constexpr uint32_t LBO128 = 128 * 16;
    const uint32_t xo_a = smem_u32(smem + sm.xo), obf_a = smem_u32(smem + sm.obf);
    const uint32_t wq_a = smem_u32(smem + sm.wq), wo_a = smem_u32(smem + sm.wo);
    uint32_t cs = 0;  // scratch passes so far
    if (lead) mbar_wait(w_full, 0);
    // One scratch pass: C[128 x N] = A[128 x DP] . W[n0.., :]^T into the scratch columns.  The group barrier makes
    // the previous pass's reads (and this pass's operand writes) complete before the lead warp issues.
    auto scratch_pass = [&](uint32_t a_addr, uint32_t w_addr, int w_rows, int n0, int N, uint64_t* also_commit) {
      tc_fence_before();
      named_bar_sync(1, 128);
      if (lead) {
        tc_fence_after();
        umma_ss_chain<DP / 16>(tmem_base + COL_SCR, make_desc(a_addr, LBO128, 128), (2 * LBO128) >> 4,
                               make_desc(w_addr + (n0 >> 3) * 128, w_rows * 16, 128), (2 * w_rows * 16) >> 4,
                               make_idesc(TILE_M, N), 0);
        umma_commit(scr_full);
        if (also_commit) umma_commit(also_commit);
      }
      mbar_wait(scr_full, cs & 1);
      ++cs;
      tc_fence_after();
    };
    // 16 fp32 accumulator columns -> two bf16 k-groups of the Q operand
    auto store_q16 = [&](const uint32_t (&r)[16], uint8_t* dst) {
      uint4 q0, q1;
      q0.x = pack_op(__uint_as_float(r[0]), __uint_as_float(r[1]));
      q0.y = pack_op(__uint_as_float(r[2]), __uint_as_float(r[3]));
      q0.z = pack_op(__uint_as_float(r[4]), __uint_as_float(r[5]));
      q0.w = pack_op(__uint_as_float(r[6]), __uint_as_float(r[7]));
      q1.x = pack_op(__uint_as_float(r[8]), __uint_as_float(r[9]));
      q1.y = pack_op(__uint_as_float(r[10]), __uint_as_float(r[11]));
      q1.z = pack_op(__uint_as_float(r[12]), __uint_as_float(r[13]));
      q1.w = pack_op(__uint_as_float(r[14]), __uint_as_float(r[15]));
      *reinterpret_cast<uint4*>(dst) = q0;
      *reinterpret_cast<uint4*>(dst + TILE_M * 16) = q1;
    };
    // stage X of local tile t as the bf16 A operand, project it, and turn the result into bf16 Q tiles
    auto prepare = [&](uint32_t t) {
      const uint32_t tb = t & 1;
      int64_t row0, sample;
      int rows;
      tile_rows(t, row0, rows, sample);
      const bool valid = row < rows;
      const float4* src = reinterpret_cast<const float4*>(smem + sm.x32 + tb * sm.tile_stride) + (size_t)row * (d / 4);
      mbar_wait_relaxed(&t_full[tb], (t >> 1) & 1, 200);
      FFD_STAMP(warp == 8 && t == 3, 404);
      constexpr int NG = DP / 8;
#pragma unroll 5
      for (int kg = 0; kg < NG; ++kg) {
        const int c = kg * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
        if (c < d) {
          if (valid) a = src[2 * kg];
        } else if (c == d) {
          a.x = 1.f, a.y = 1.f;
        }
        if (c + 4 < d) {
          if (valid) b4 = src[2 * kg + 1];
        } else if (c + 4 == d) {
          b4.x = 1.f, b4.y = 1.f;
        }
        uint4 q;
        q.x = pack_op(a.x, a.y);
        q.y = pack_op(a.z, a.w);
        q.z = pack_op(b4.x, b4.y);
        q.w = pack_op(b4.z, b4.w);
        *reinterpret_cast<uint4*>(xo + row_off + kg * (TILE_M * 16)) = q;
      }
      fence_proxy_async();
      FFD_STAMP(warp == 8 && t == 3, 403);
      uint8_t* qb = smem + sm.qbf + tb * sm.qbf_bytes;
      for (int n0 = 0; n0 < QN; n0 += SCR_COLS) {
        const int N = QN - n0 < SCR_COLS ? QN - n0 : SCR_COLS;
        scratch_pass(xo_a, wq_a, QN, n0, N, nullptr);
        // two 16-column units in flight per TMEM round trip
        for (int u = 0; u < N; u += 32) {
          uint32_t ra[16], rb[16];
          const bool two = u + 16 < N;
          FFD_TMEM_LD16(scr + u, ra);
          if (two) FFD_TMEM_LD16(scr + u + 16, rb);
          tmem_ld_wait();
          const int kg = (n0 + u) >> 3;
          store_q16(ra, qb + row_off + kg * (TILE_M * 16));
          if (two) store_q16(rb, qb + row_off + (kg + 2) * (TILE_M * 16));
        }
      }
      fence_proxy_async();
      mbar_arrive(&qbf_full[tb]);
    };
    // out-projection of local tile t (A operand: the normalised O columns written by the read-out warps) +
    // residual -> LayerNorm, in place in the fp32 tile buffer (one thread owns a row)
    auto finish = [&](uint32_t t) {
      const uint32_t tb = t & 1;
      float4* myrow = reinterpret_cast<float4*>(smem + sm.x32 + tb * sm.tile_stride) + (size_t)row * (d / 4);
      float s1 = 0.f, s2 = 0.f;
      if (lead) mbar_wait_relaxed(obf_full, t & 1, 100);
      for (int n0 = 0; n0 < DP; n0 += SCR_COLS) {
        const int N = DP - n0 < SCR_COLS ? DP - n0 : SCR_COLS;
        scratch_pass(obf_a, wo_a, DP, n0, N, n0 + SCR_COLS >= DP ? obf_empty : nullptr);
        for (int u = 0; u < N; u += 16) {
          uint32_t r[16];
          FFD_TMEM_LD16(scr + u, r);
          float4 res[4];  // residual: issued before the TMEM round trip completes
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = n0 + u + 4 * j;
            res[j] = c < d ? myrow[c >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = n0 + u + 4 * j;
            if (c < d) {
              float4 v = res[j];
              v.x += __uint_as_float(r[4 * j]);
              v.y += __uint_as_float(r[4 * j + 1]);
              v.z += __uint_as_float(r[4 * j + 2]);
              v.w += __uint_as_float(r[4 * j + 3]);
              myrow[c >> 2] = v;
              s1 += (v.x + v.y) + (v.z + v.w);
              s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
            }
          }
        }
      }
      const float mean = s1 / (float)d;
      const float rstd = rsqrtf(fmaxf(s2 / (float)d - mean * mean, 0.f) + eps);
#pragma unroll 6
      for (int c = 0; c < d; c += 4) {
        const float4 v = myrow[c >> 2];
        float4 o;
        o.x = (v.x - mean) * rstd * consts[c] + consts[DP + c];
        o.y = (v.y - mean) * rstd * consts[c + 1] + consts[DP + c + 1];
        o.z = (v.z - mean) * rstd * consts[c + 2] + consts[DP + c + 2];
        o.w = (v.w - mean) * rstd * consts[c + 3] + consts[DP + c + 3];
        myrow[c >> 2] = o;
      }
      fence_proxy_async();
      mbar_arrive(&y_ready[tb]);
    };
    prepare(0);
    for (uint32_t t = 0; t < my_tiles; ++t) {
      FFD_STAMP(warp == 8 && t == 2, 400);
      if (t >= 1) finish(t - 1);
      FFD_STAMP(warp == 8 && t == 2, 401);
      if (t + 1 < my_tiles) prepare(t + 1);
      FFD_STAMP(warp == 8 && t == 2, 402);
    }
    finish(my_tiles - 1);
    if (lead) mbar_wait_relaxed(obf_empty, (my_tiles - 1) & 1, 200);  // drain
  } else {
    // ================= softmax groups: heads and nothing else =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int wg = (warp - 12) >> 2;       // 0..3: quarter of the key columns
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t scol = tmem_base + lane_base + COL_S + wg * NQ;  // this group's score columns
    const uint32_t pcol0 = tmem_base + lane_base + COL_P0 + wg * NP;  // ... and its packed P columns
    // KQ + 1 groups of 16 score registers: the current head's KQ chunks and one spare group, so that ALL of the
    // next head's scores can be loaded (and the score buffer released) while this head's last chunk is still in
    // use.  The last chunk and the spare group swap roles from one head to the next (loop unrolled by two).
    uint32_t r[16 * (KQ + 1)];
    auto chunk_max = [&](int j, float m) {
#pragma unroll
      for (int e = 0; e < 16; e += 2) m = max3(m, __uint_as_float(r[16 * j + e]), __uint_as_float(r[16 * j + e + 1]));
      return m;
    };
    // The four softmax warps of a sub-partition (one per group) do identical work; started together they stay in
    // lockstep and their non-MUFU sections (row maximum, TMEM stores, barrier traffic) leave the MUFU pipe idle
    // together.  Start them a quarter of a head apart.
    mbar_wait(s_full, 0);
    __nanosleep(wg * FFD_STAGGER_NS);
    tc_fence_after();
#pragma unroll
    for (int j = 0; j < KQ; ++j) FFD_TMEM_LD16(scol + 16 * j, (r + 16 * j));
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(s_empty);
    float mb = -INFINITY;
#pragma unroll
    for (int j = 0; j < KQ; ++j) mb = chunk_max(j, mb);
    // p = exp2(s - m_q) as fp16 pairs with m_q the row maximum over THIS group's quarter of the keys: 16 score
    // columns -> 8 packed columns.  Padded keys need no mask: slot 15 of every Q tile is the constant 1 and slot
    // 15 of a padded K row is -29952 (a quarter made only of padding gets m_q = -29952 and weight
    // exp2(m_q - max) = 0 in the combine).
#define FFD_EXP16(G_, mc_, pk_)                                                                   \
  _Pragma("unroll") for (int e = 0; e < 16; e += 2)                                               \
      pk_[e >> 1] = ((POLY_PAIRS >> (e >> 1)) & 1)                                                   \
                        ? ex2_pack_f16_shift_poly(__uint_as_float(r[16 * (G_) + e]), __uint_as_float(r[16 * (G_) + e + 1]), (mc_)) \
                        : ex2_pack_f16_shift(__uint_as_float(r[16 * (G_) + e]), __uint_as_float(r[16 * (G_) + e + 1]), (mc_))
#define FFD_SOFTMAX_HEAD(LAST, SPARE, n_, B_)                                                      \
  {                                                                                               \
    const uint32_t n = (n_);                                                                      \
    constexpr uint32_t b = (B_); /* = n & 1: the loop is unrolled by two, so buffers and parities are constants */ \
    const uint32_t pcol = pcol0 + b * (COL_P1 - COL_P0);                                          \
    const bool has_next = n + 1 < n_heads;                                                        \
    const float mc = -mb;                                                                         \
    mq[((n & 3) * 4 + wg) * TILE_M + row] = mb;                                                   \
    FFD_STAMP(warp == 12 && n - 2u * NV < (uint32_t)NV, 128 + (n - 2u * NV) * 4 + 0);                                 \
    uint32_t pk[8];                                                                               \
    _Pragma("unroll") for (int j = 0; j < KQ - 1; ++j) {                                          \
      FFD_EXP16(j, mc, pk);                                                                       \
      if (j == 0) {                                                                               \
        mbar_wait(&p_empty[b], ((n >> 1) & 1) ^ 1); /* the P.V MMA two heads back has read this buffer */ \
        tc_fence_after();                                                                         \
      }                                                                                           \
      FFD_TMEM_ST8(pcol + 8 * j, pk);                                                             \
    }                                                                                             \
    if (has_next) {                                                                               \
      FFD_STAMP(warp == 12 && n - 2u * NV < (uint32_t)NV, 128 + (n - 2u * NV) * 4 + 1);                               \
      mbar_wait_relaxed(s_full, b ^ 1, 32); /* parity (n + 1) & 1 */                              \
      tc_fence_after();                                                                           \
      FFD_STAMP(warp == 12 && n - 2u * NV < (uint32_t)NV, 128 + (n - 2u * NV) * 4 + 2);                               \
      _Pragma("unroll") for (int j = 0; j < KQ - 1; ++j) FFD_TMEM_LD16(scol + 16 * j, (r + 16 * j)); \
      FFD_TMEM_LD16(scol + 16 * (KQ - 1), (r + 16 * (SPARE)));                                    \
    }                                                                                             \
    /* first half of the last chunk's exponentials, then (same basic block, so that ptxas interleaves them) */ \
    /* the second half together with the row maximum of the next head's scores                              */ \
    _Pragma("unroll") for (int e = 0; e < 8; e += 2)                                              \
        pk[e >> 1] = ex2_pack_f16_shift(__uint_as_float(r[16 * (LAST) + e]), __uint_as_float(r[16 * (LAST) + e + 1]), mc); \
    if (has_next) {                                                                               \
      tmem_ld_wait();                                                                             \
      tc_fence_before();                                                                          \
      mbar_arrive(s_empty); /* the score buffer may be overwritten */                             \
    }                                                                                             \
    _Pragma("unroll") for (int e = 8; e < 16; e += 2)                                             \
        pk[e >> 1] = ex2_pack_f16_shift(__uint_as_float(r[16 * (LAST) + e]), __uint_as_float(r[16 * (LAST) + e + 1]), mc); \
    { /* unconditional (stale registers after the last head: unused), to stay in the same basic block */ \
      float mj[KQ]; /* independent chains: the reduction is latency-bound otherwise */            \
      _Pragma("unroll") for (int j = 0; j < KQ - 1; ++j) mj[j] = chunk_max(j, -INFINITY);         \
      mj[KQ - 1] = chunk_max(SPARE, -INFINITY);                                                   \
      mb = mj[0];                                                                                 \
      _Pragma("unroll") for (int j = 1; j < KQ; ++j) mb = fmaxf(mb, mj[j]);                       \
    }                                                                                             \
    if (KQ == 1) {                                                                                \
      mbar_wait(&p_empty[b], ((n >> 1) & 1) ^ 1);                                                 \
      tc_fence_after();                                                                           \
    }                                                                                             \
    FFD_TMEM_ST8(pcol + 8 * (KQ - 1), pk);                                                        \
    tmem_st_wait();                                                                               \
    tc_fence_before();                                                                            \
    mbar_arrive(&p_full[b]);                                                                      \
    FFD_STAMP(warp == 12 && n - 2u * NV < (uint32_t)NV, 128 + (n - 2u * NV) * 4 + 3);                                 \
  }
    for (uint32_t n2 = 0; n2 < n_heads; n2 += 2) {
      FFD_SOFTMAX_HEAD(KQ - 1, KQ, n2, 0);
      if (n2 + 1 < n_heads) FFD_SOFTMAX_HEAD(KQ, KQ - 1, n2 + 1, 1);
    }
#undef FFD_SOFTMAX_HEAD
#undef FFD_EXP16
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#undef FFD_STAMP
}

// ---- v12 (EXPERIMENT, compiled only with -DFFD_BUILD_ATTN_QUARTERS): four independent quarter streams ------------
// Timeline of v11 (tools/attn_stamps.py): every softmax warp spends ~1540 of ~2050 cycles per head in its exponential
// phases; all sixteen move in lock-step through one score buffer and one P barrier, so the MUFU pipe idles the other
// 25 % for everybody at once.  Here every quarter of the keys is its own pipeline: its own score MMA (N = 16 KQ) into
// one of its own two score buffers, P written in place over the scores, its own P.V chain, barriers and MMA-issuing
// warp (1024 threads).  TMEM: S/P 4 x 2 x 16 KQ, O 64, scratch 64 = 512 columns.  Results equal v11's to 2.5e-5.
// Measured (ECG, 16 384 samples, attention block incl. image pack): v11 2.02 ms; v12 2.22 ms; v12 with the
// exponential phase passed round the groups by named barriers (FFD_ATTN_RING) 2.33 ms (two sets of two groups) and
// 2.91 ms (one group at a time).  Why it loses: one warp alone needs ~750 cycles for the 42 MUFU + 3 polynomial pairs
// of a head (8 cycles per MUFU would be 336) -- the exponential phase is latency-bound per warp, and it is exactly the
// four-warps-in-step interleaving of v11 that keeps the pipe at 75 %; left alone the four streams fall back into
// step anyway (they re-align at every tile boundary) and then expose the un-prefetched TMEM load + max (~550 cycles)
// that v11's software pipeline hides.  Kept as the record of the experiment; not in the library.
#ifdef FFD_BUILD_ATTN_QUARTERS
template <int DP, int KQ>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_quarters_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ wqimg,
                      const uint8_t* __restrict__ woimg, const uint8_t* __restrict__ kvimg,
                      const float* __restrict__ lnw, const float* __restrict__ lnb, int64_t M, int d, int H, int hd,
                      int G, int QN, int NCH, int64_t n_tiles, float eps, int Lown, int tps,
                      long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int CK = 64 * KQ;   // keys per chunk
  constexpr int NQ = 16 * KQ;   // key columns per softmax group
  constexpr int NP = 8 * KQ;    // packed P columns per softmax group
  const int NV = H * NCH;       // virtual heads per tile
  const AttnSmem sm = attn_smem(DP, QN, CK, d);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* k_full = bars;                       // [K_STAGES] tx
  uint64_t* k_empty = k_full + K_STAGES;         // [K_STAGES] commit
  uint64_t* v_full = k_empty + K_STAGES;         // [V_STAGES] tx
  uint64_t* v_empty = v_full + V_STAGES;         // [V_STAGES] commit
  uint64_t* w_full = v_empty + V_STAGES;         // [1] tx (Wq + Wo, once)
  uint64_t* t_full = w_full + 1;                 // [2] tx: fp32 tile landed
  uint64_t* y_ready = t_full + 2;                // [2] 128 arrivals: LayerNorm output written to the tile buffer
  uint64_t* scr_full = y_ready + 2;              // [1] commit: a scratch pass has been computed
  uint64_t* qbf_full = scr_full + 1;             // [2] 128 arrivals: bf16 Q tiles of a token tile written
  uint64_t* s_full = qbf_full + 2;               // [4][2] commit: scores of quarter q of a virtual head (buffer n & 1)
  uint64_t* p_full = s_full + 8;                 // [4][2] 128 arrivals: quarter q's P written in place over its scores
  uint64_t* o_full = p_full + 8;                 // [1] 4 commits: the four quarter accumulators of a virtual head
  uint64_t* o_empty = o_full + 1;                // [1] 128 arrivals (read-out warps)
  uint64_t* obf_full = o_empty + 1;              // [1] 128 arrivals (read-out warps)
  uint64_t* obf_empty = obf_full + 1;            // [1] commit: the out-projection has read its operand
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.tmem);
  float* consts = reinterpret_cast<float*>(smem + sm.consts);
  float* mq = reinterpret_cast<float*>(smem + sm.mq);   // [4][4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t my_tiles = (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // >= 1
  // Rows of this CTA's t-th tile.  Lown == 0: K/V are the shared cache, tiles are consecutive 128-row blocks of
  // [0, M).  Lown > 0 (stock attention, cached_transformer.py:123-134 / nn.MultiheadAttention): every sample
  // attends to its OWN keys, so tiles are aligned to samples (tps tiles of <= 128 rows per sample of Lown rows)
  // and the K/V images are per sample.
  auto tile_rows = [&](uint32_t t, int64_t& row0, int& rows, int64_t& sample) {
    const int64_t gt = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
    if (Lown == 0) {
      sample = 0;
      row0 = gt * TILE_M;
      rows = (int)(M - row0 < TILE_M ? M - row0 : TILE_M);
    } else {
      sample = gt / tps;
      const int part = (int)(gt - sample * tps);
      row0 = sample * Lown + (int64_t)part * TILE_M;
      rows = Lown - part * TILE_M < TILE_M ? Lown - part * TILE_M : TILE_M;
    }
  };
  const uint32_t n_heads = my_tiles * (uint32_t)NV;  // virtual heads this CTA runs through
#ifdef FFD_ATTN_STAMPS  // timeline of local tile 2 of CTA 0 (tools/attn_stamps.py)
#define FFD_STAMP(cond, slot) do { if (dbg && blockIdx.x == 0 && lane == 0 && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define FFD_STAMP(cond, slot) do { } while (0)
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < K_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 4);  // the four quarter streams each commit once per image
    }
    for (int i = 0; i < V_STAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 4);
    }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&t_full[b], 1);
      mbar_init(&y_ready[b], 128);
      mbar_init(&qbf_full[b], 128);
    }
    for (int q = 0; q < 8; ++q) {
      mbar_init(&s_full[q], 1);
      mbar_init(&p_full[q], 128);
    }
    mbar_init(scr_full, 1);
    mbar_init(o_full, 4);
    mbar_init(o_empty, 128);
    mbar_init(obf_full, 128);
    mbar_init(obf_empty, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * DP; i += ATT2_THREADS) {
    const int w = i / DP, c = i % DP;
    consts[i] = c < d ? (w == 0 ? lnw[c] : lnb[c]) : 0.f;
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // provably warp-uniform

  // register budget per warpgroup (896 threads x 72 at launch): the softmax warps hold two heads' scores
  // (set at the top of each role branch so that ptxas allocates per branch)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ================= K/V image producer =================
    if (lane == 0) {
      mbar_expect_tx(w_full, sm.wq_bytes + sm.wo_bytes);
      bulk_load(smem + sm.wq, wqimg, sm.wq_bytes, w_full);
      bulk_load(smem + sm.wo, woimg, sm.wo_bytes, w_full);
      uint32_t nk = 0, nv = 0;
      const size_t img_stride = (size_t)sm.k_bytes + sm.v_bytes;
      const size_t sample_stride = Lown ? (size_t)NV * img_stride : 0;
      auto load_k = [&](const uint8_t* base, uint32_t v) {
        const uint32_t st = nk % K_STAGES, u = nk / K_STAGES;
        mbar_wait_relaxed(&k_empty[st], (u & 1) ^ 1, 200);
        mbar_expect_tx(&k_full[st], sm.k_bytes);
        bulk_load(smem + sm.kring + st * sm.k_bytes, base + v * img_stride, sm.k_bytes, &k_full[st]);
        ++nk;
      };
      auto load_v = [&](const uint8_t* base, uint32_t v) {
        const uint32_t st = nv % V_STAGES, u = nv / V_STAGES;
        mbar_wait_relaxed(&v_empty[st], (u & 1) ^ 1, 200);
        mbar_expect_tx(&v_full[st], sm.v_bytes);
        bulk_load(smem + sm.vring + st * sm.v_bytes, base + v * img_stride + sm.k_bytes, sm.v_bytes, &v_full[st]);
        ++nv;
      };
      auto tile_base = [&](uint32_t t) {
        int64_t row0, sample;
        int rows;
        tile_rows(t, row0, rows, sample);
        return kvimg + (size_t)sample * sample_stride;
      };
      // K images run two virtual heads ahead of the V images (the score MMAs run ahead of the P.V MMAs)
      uint32_t vk = 0, vv = 0, tk = 0, tv = 0;  // virtual head inside its tile / tile of the next K and V image
      const uint8_t* kbase = tile_base(0);
      const uint8_t* vbase = kbase;
      auto next_k = [&]() {
        load_k(kbase, vk);
        if (++vk == (uint32_t)NV) {
          vk = 0;
          if (++tk < my_tiles) kbase = tile_base(tk);
        }
      };
      next_k();
      if (n_heads > 1) next_k();
      for (uint32_t n = 0; n < n_heads; ++n) {
        load_v(vbase, vv);
        if (++vv == (uint32_t)NV) {
          vv = 0;
          if (++tv < my_tiles) vbase = tile_base(tv);
        }
        if (n + 2 < n_heads) next_k();
      }
    }
  } else if (warp == 2) {
    // ================= token-tile load / store =================
    if (lane == 0) {
      auto load_tile = [&](uint32_t t) {
        const uint32_t b = t & 1;
        int64_t row0, sample;
        int rows;
        tile_rows(t, row0, rows, sample);
        mbar_expect_tx(&t_full[b], (uint32_t)(rows * d * 4));
        bulk_load(smem + sm.x32 + b * sm.tile_stride, x + row0 * d, (uint32_t)(rows * d * 4), &t_full[b]);
      };
      load_tile(0);
      if (my_tiles > 1) load_tile(1);
      for (uint32_t t = 0; t < my_tiles; ++t) {
        const uint32_t b = t & 1;
        int64_t row0, sample;
        int rows;
        tile_rows(t, row0, rows, sample);
        mbar_wait_relaxed(&y_ready[b], (t >> 1) & 1, 100);
        bulk_store(y + row0 * d, smem + sm.x32 + b * sm.tile_stride, (uint32_t)(rows * d * 4));
        bulk_commit();
        bulk_wait_read();
        if (t + 2 < my_tiles) load_tile(t + 2);  // the buffer is free again
      }
      bulk_wait_all();
    }
  }
  } else if (warp >= 28) {
    // ================= MMA issuers (whole warp converged, one elected lane issues): warp 28 + q serves quarter stream q.
    // A stream's step: P.V of its quarter of virtual head n into its own accumulator, then -- behind it on the in-order
    // tensor pipe, into the very columns P_n is read from -- the scores of its quarter of virtual head n + 2.  The
    // scores of head n + 1 sit in the stream's other buffer already, so a softmax group never waits for an MMA round
    // trip; the streams are independent up to the shared O read-out (one head) and the K/V image rings, so the four
    // groups drift out of phase and the MUFU pipe of every sub-partition always finds one of its four warps in its
    // exponential phase (v11: all sixteen warps leave it together, 25 % of every head).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int q = warp - 28;
    const uint32_t idesc_s = make_idesc(TILE_M, NQ);
    constexpr uint32_t IDESC_PV = make_idesc_f16(TILE_M, HP);
    constexpr uint32_t LBO128 = 128 * 16;
    constexpr uint32_t VSTEP = (2 * HP * 16) >> 4;  // 16 keys of the V image, in 16-byte units
    const uint32_t qbf = smem_u32(smem + sm.qbf);
    const uint32_t kring = smem_u32(smem + sm.kring);
    const uint32_t vring = smem_u32(smem + sm.vring);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](const uint64_t* b) { return bar0 + (uint32_t)((b - bars) * 8); };
    // incremental position of the NEXT score MMA of this stream: virtual head number, tile, head-in-tile counters
    uint32_t s_n = 0, s_t = 0, s_v = 0, s_c = 0, s_g = 0, s_qt = 0;
    auto s_advance = [&]() {
      ++s_n;
      if (++s_c == (uint32_t)NCH) {
        s_c = 0;
        if (++s_g == (uint32_t)G) s_g = 0, ++s_qt;
      }
      if (++s_v == (uint32_t)NV) s_v = 0, s_c = 0, s_g = 0, s_qt = 0, ++s_t;
    };
    auto issue_s = [&]() {  // scores of this quarter of virtual head s_n into buffer s_n & 1 (K image and Q tiles are there)
      const uint32_t st = s_n % K_STAGES, bf = s_n & 1;
      umma(tmem_base + COL_S + NQ * (2 * q + bf), make_desc(qbf + (s_t & 1) * sm.qbf_bytes + s_qt * 2 * LBO128, LBO128, 128),
           make_desc(kring + st * sm.k_bytes + (uint32_t)(NQ * q / 8) * 128, CK * 16, 128), idesc_s, 0);
      umma_commit(&s_full[2 * q + bf]);
      umma_commit(&k_empty[st]);
      s_advance();
    };
    // prologue: the first two virtual heads' scores
    for (uint32_t i = 0; i < 2 && i < n_heads; ++i) {
      if (s_v == 0) mbar_wait_relaxed(&qbf_full[s_t & 1], (s_t >> 1) & 1, 60);
      mbar_wait(&k_full[s_n % K_STAGES], (s_n / K_STAGES) & 1);
      tc_fence_after();
      issue_s();
    }
    for (uint32_t n = 0; n < n_heads; ++n) {
      const uint32_t bf = n & 1, use = n >> 1;
      const uint32_t stv = n % V_STAGES, stk = s_n % K_STAGES;
      const bool more = s_n < n_heads;  // s_n == n + 2 here
      // all the barriers of this step in one round trip: lane 0 this quarter's P_n complete, 1 the previous virtual head's
      // O read out, 2 V image landed, 3 K image of virtual head n + 2 landed, 4 its tile's Q tiles written
      {
        uint32_t wb = bar(&p_full[2 * q + bf]), wp = use & 1;
        if (lane == 1) wb = bar(o_empty), wp = (n & 1) ^ 1;
        if (lane == 2) wb = bar(&v_full[stv]), wp = (n / V_STAGES) & 1;
        if (lane == 3 && more) wb = bar(&k_full[stk]), wp = (s_n / K_STAGES) & 1;
        if (lane == 4 && more && s_v == 0) wb = bar(&qbf_full[s_t & 1]), wp = (s_t >> 1) & 1;
        FFD_STAMP(n - 2u * NV < (uint32_t)NV, 1024 + (q * 32 + (n - 2u * NV)) * 4 + 0);
        mbar_wait_lanes(wb, wp);
      }
      tc_fence_after();
      FFD_STAMP(n - 2u * NV < (uint32_t)NV, 1024 + (q * 32 + (n - 2u * NV)) * 4 + 1);
      umma_ts_chain<KQ>(tmem_base + COL_O + 16 * q, tmem_base + COL_S + NQ * (2 * q + bf), 8,
                        make_desc(vring + stv * sm.v_bytes, HP * 16, 128) + (uint64_t)(q * KQ) * VSTEP, VSTEP, IDESC_PV, 0);
      umma_commit(o_full);
      umma_commit(&v_empty[stv]);
      FFD_STAMP(n - 2u * NV < (uint32_t)NV, 1024 + (q * 32 + (n - 2u * NV)) * 4 + 2);
      if (more) issue_s();
      FFD_STAMP(n - 2u * NV < (uint32_t)NV, 1024 + (q * 32 + (n - 2u * NV)) * 4 + 3);
    }
    mbar_wait_relaxed(o_full, (n_heads - 1) & 1, 200);  // drain: nothing of this warp is in flight when the CTA exits
  } else if (warp < 8) {
    // ================= O read-out warps =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    uint8_t* ob = smem + sm.obf;
    // constant columns of the out-proj operand: d, d+1 = 1 (bias hi/lo), the rest of the padding = 0
    for (int c = d; c < DP; ++c)
      *reinterpret_cast<op16*>(ob + row_off + (c >> 3) * (TILE_M * 16) + (c & 7) * 2) =
          to_op((c == d || c == d + 1) ? 1.f : 0.f);
    uint32_t n = 0;
    for (uint32_t t = 0; t < my_tiles; ++t) {
      for (int h = 0; h < H; ++h) {
        // With hd <= 7 the row sum sits in slot 7 (pack_kv_kernel), so the low 8 columns of every quarter
        // accumulator hold everything: ONE TMEM round trip per virtual head.  (The P.V MMA of the next head waits
        // for this read: the round trip is on the kernel's critical loop.)
        if (hd <= 7) {
          float acc[8];
          float mrun = 0.f;
          for (int c = 0; c < NCH; ++c, ++n) {
            mbar_wait_relaxed(o_full, n & 1, 32);
            tc_fence_after();
            FFD_STAMP(warp == 4 && t == 2, 256 + (h * NCH + c) * 2 + 0);
            uint32_t r[32];
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O, r);
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O + 16, (r + 8));
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O + 32, (r + 16));
            FFD_TMEM_LD8(tmem_base + lane_base + COL_O + 48, (r + 24));
            // combine the four quarters: weights exp2(m_q - max m_q)
            const float* mh = mq + ((n & 3) * 4) * TILE_M + row;
            const float q0 = mh[0], q1 = mh[TILE_M], q2 = mh[2 * TILE_M], q3 = mh[3 * TILE_M];
            float mm = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
            if (c > 0) mm = fmaxf(mm, mrun);
            const float w0 = exp2_poly3(q0 - mm), w1 = exp2_poly3(q1 - mm), w2 = exp2_poly3(q2 - mm),
                        w3 = exp2_poly3(q3 - mm);
            const float wp = c > 0 ? exp2_poly3(mrun - mm) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = c > 0 ? acc[j] * wp : 0.f;
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty);
            FFD_STAMP(warp == 4 && t == 2, 256 + (h * NCH + c) * 2 + 1);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w0, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[8 + j]), __uint_as_float(r[9 + j]), w1, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[16 + j]), __uint_as_float(r[17 + j]), w2, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[24 + j]), __uint_as_float(r[25 + j]), w3, acc[j], acc[j + 1]);
            }
            mrun = mm;
          }
          const float inv = rcp_fma(acc[7]);
          // the out-projection of the previous tile must have read its operand before the first overwrite
          if (h == 0 && t >= 1) mbar_wait_relaxed(obf_empty, (t - 1) & 1, 60);
          if ((hd & 1) == 0) {  // even head dim: the head's columns start even, two columns per store
#pragma unroll
            for (int j = 0; j < 6; j += 2) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<uint32_t*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    pack_op(acc[j] * inv, acc[j + 1] * inv);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<op16*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    to_op(acc[j] * inv);
              }
            }
          }
        } else {
          float acc[HP];
          float mrun = 0.f;
          for (int c = 0; c < NCH; ++c, ++n) {
            mbar_wait_relaxed(o_full, n & 1, 32);
            tc_fence_after();
            uint32_t r[32];
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O, r);
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 16, (r + 16));
            const float* mh = mq + ((n & 3) * 4) * TILE_M + row;
            const float q0 = mh[0], q1 = mh[TILE_M], q2 = mh[2 * TILE_M], q3 = mh[3 * TILE_M];
            float mm = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
            if (c > 0) mm = fmaxf(mm, mrun);
            const float w0 = exp2_poly3(q0 - mm), w1 = exp2_poly3(q1 - mm), w2 = exp2_poly3(q2 - mm),
                        w3 = exp2_poly3(q3 - mm);
            const float wp = c > 0 ? exp2_poly3(mrun - mm) : 0.f;
#pragma unroll
            for (int j = 0; j < HP; ++j) acc[j] = c > 0 ? acc[j] * wp : 0.f;
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
              fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w0, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[16 + j]), __uint_as_float(r[17 + j]), w1, acc[j], acc[j + 1]);
            }
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 32, r);
            FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 48, (r + 16));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty);
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
              fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w2, acc[j], acc[j + 1]);
              fma2(acc[j], acc[j + 1], __uint_as_float(r[16 + j]), __uint_as_float(r[17 + j]), w3, acc[j], acc[j + 1]);
            }
            mrun = mm;
          }
          const float inv = rcp_fma(acc[HP - 1]);
          if (h == 0 && t >= 1) mbar_wait_relaxed(obf_empty, (t - 1) & 1, 60);
          if ((hd & 1) == 0) {
#pragma unroll
            for (int j = 0; j < HP - 2; j += 2) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<uint32_t*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    pack_op(acc[j] * inv, acc[j + 1] * inv);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < HP - 1; ++j) {
              if (j < hd) {
                const int col = h * hd + j;
                *reinterpret_cast<op16*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                    to_op(acc[j] * inv);
              }
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(obf_full);
    }
  } else if (warp < 12) {
    // ================= tile warps: prepare tile t+1 / finish tile t-1 while the heads of tile t run ==========
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    const uint32_t scr = tmem_base + lane_base + COL_SCR;
    uint8_t* xo = smem + sm.xo;
    const bool lead = warp == 8;  // issues the scratch MMAs of the group (whole warp converged)
    constexpr uint32_t LBO128 = 128 * 16;
    const uint32_t xo_a = smem_u32(smem + sm.xo), obf_a = smem_u32(smem + sm.obf);
    const uint32_t wq_a = smem_u32(smem + sm.wq), wo_a = smem_u32(smem + sm.wo);
    uint32_t cs = 0;  // scratch passes so far
    if (lead) mbar_wait(w_full, 0);
    // One scratch pass: C[128 x N] = A[128 x DP] . W[n0.., :]^T into the scratch columns.  The group barrier makes
    // the previous pass's reads (and this pass's operand writes) complete before the lead warp issues.
    auto scratch_pass = [&](uint32_t a_addr, uint32_t w_addr, int w_rows, int n0, int N, uint64_t* also_commit) {
      tc_fence_before();
      named_bar_sync(1, 128);
      if (lead) {
        tc_fence_after();
        umma_ss_chain<DP / 16>(tmem_base + COL_SCR, make_desc(a_addr, LBO128, 128), (2 * LBO128) >> 4,
                               make_desc(w_addr + (n0 >> 3) * 128, w_rows * 16, 128), (2 * w_rows * 16) >> 4,
                               make_idesc(TILE_M, N), 0);
        umma_commit(scr_full);
        if (also_commit) umma_commit(also_commit);
      }
      mbar_wait(scr_full, cs & 1);
      ++cs;
      tc_fence_after();
    };
    // 16 fp32 accumulator columns -> two bf16 k-groups of the Q operand
    auto store_q16 = [&](const uint32_t (&r)[16], uint8_t* dst) {
      uint4 q0, q1;
      q0.x = pack_op(__uint_as_float(r[0]), __uint_as_float(r[1]));
      q0.y = pack_op(__uint_as_float(r[2]), __uint_as_float(r[3]));
      q0.z = pack_op(__uint_as_float(r[4]), __uint_as_float(r[5]));
      q0.w = pack_op(__uint_as_float(r[6]), __uint_as_float(r[7]));
      q1.x = pack_op(__uint_as_float(r[8]), __uint_as_float(r[9]));
      q1.y = pack_op(__uint_as_float(r[10]), __uint_as_float(r[11]));
      q1.z = pack_op(__uint_as_float(r[12]), __uint_as_float(r[13]));
      q1.w = pack_op(__uint_as_float(r[14]), __uint_as_float(r[15]));
      *reinterpret_cast<uint4*>(dst) = q0;
      *reinterpret_cast<uint4*>(dst + TILE_M * 16) = q1;
    };
    // stage X of local tile t as the bf16 A operand, project it, and turn the result into bf16 Q tiles
    auto prepare = [&](uint32_t t) {
      const uint32_t tb = t & 1;
      int64_t row0, sample;
      int rows;
      tile_rows(t, row0, rows, sample);
      const bool valid = row < rows;
      const float4* src = reinterpret_cast<const float4*>(smem + sm.x32 + tb * sm.tile_stride) + (size_t)row * (d / 4);
      mbar_wait_relaxed(&t_full[tb], (t >> 1) & 1, 200);
      FFD_STAMP(warp == 8 && t == 3, 404);
      constexpr int NG = DP / 8;
#pragma unroll 5
      for (int kg = 0; kg < NG; ++kg) {
        const int c = kg * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
        if (c < d) {
          if (valid) a = src[2 * kg];
        } else if (c == d) {
          a.x = 1.f, a.y = 1.f;
        }
        if (c + 4 < d) {
          if (valid) b4 = src[2 * kg + 1];
        } else if (c + 4 == d) {
          b4.x = 1.f, b4.y = 1.f;
        }
        uint4 q;
        q.x = pack_op(a.x, a.y);
        q.y = pack_op(a.z, a.w);
        q.z = pack_op(b4.x, b4.y);
        q.w = pack_op(b4.z, b4.w);
        *reinterpret_cast<uint4*>(xo + row_off + kg * (TILE_M * 16)) = q;
      }
      fence_proxy_async();
      FFD_STAMP(warp == 8 && t == 3, 403);
      uint8_t* qb = smem + sm.qbf + tb * sm.qbf_bytes;
      for (int n0 = 0; n0 < QN; n0 += SCR_COLS) {
        const int N = QN - n0 < SCR_COLS ? QN - n0 : SCR_COLS;
        scratch_pass(xo_a, wq_a, QN, n0, N, nullptr);
        // two 16-column units in flight per TMEM round trip
        for (int u = 0; u < N; u += 32) {
          uint32_t ra[16], rb[16];
          const bool two = u + 16 < N;
          FFD_TMEM_LD16(scr + u, ra);
          if (two) FFD_TMEM_LD16(scr + u + 16, rb);
          tmem_ld_wait();
          const int kg = (n0 + u) >> 3;
          store_q16(ra, qb + row_off + kg * (TILE_M * 16));
          if (two) store_q16(rb, qb + row_off + (kg + 2) * (TILE_M * 16));
        }
      }
      fence_proxy_async();
      mbar_arrive(&qbf_full[tb]);
    };
    // out-projection of local tile t (A operand: the normalised O columns written by the read-out warps) +
    // residual -> LayerNorm, in place in the fp32 tile buffer (one thread owns a row)
    auto finish = [&](uint32_t t) {
      const uint32_t tb = t & 1;
      float4* myrow = reinterpret_cast<float4*>(smem + sm.x32 + tb * sm.tile_stride) + (size_t)row * (d / 4);
      float s1 = 0.f, s2 = 0.f;
      if (lead) mbar_wait_relaxed(obf_full, t & 1, 100);
      for (int n0 = 0; n0 < DP; n0 += SCR_COLS) {
        const int N = DP - n0 < SCR_COLS ? DP - n0 : SCR_COLS;
        scratch_pass(obf_a, wo_a, DP, n0, N, n0 + SCR_COLS >= DP ? obf_empty : nullptr);
        for (int u = 0; u < N; u += 16) {
          uint32_t r[16];
          FFD_TMEM_LD16(scr + u, r);
          float4 res[4];  // residual: issued before the TMEM round trip completes
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = n0 + u + 4 * j;
            res[j] = c < d ? myrow[c >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = n0 + u + 4 * j;
            if (c < d) {
              float4 v = res[j];
              v.x += __uint_as_float(r[4 * j]);
              v.y += __uint_as_float(r[4 * j + 1]);
              v.z += __uint_as_float(r[4 * j + 2]);
              v.w += __uint_as_float(r[4 * j + 3]);
              myrow[c >> 2] = v;
              s1 += (v.x + v.y) + (v.z + v.w);
              s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
            }
          }
        }
      }
      const float mean = s1 / (float)d;
      const float rstd = rsqrtf(fmaxf(s2 / (float)d - mean * mean, 0.f) + eps);
#pragma unroll 6
      for (int c = 0; c < d; c += 4) {
        const float4 v = myrow[c >> 2];
        float4 o;
        o.x = (v.x - mean) * rstd * consts[c] + consts[DP + c];
        o.y = (v.y - mean) * rstd * consts[c + 1] + consts[DP + c + 1];
        o.z = (v.z - mean) * rstd * consts[c + 2] + consts[DP + c + 2];
        o.w = (v.w - mean) * rstd * consts[c + 3] + consts[DP + c + 3];
        myrow[c >> 2] = o;
      }
      fence_proxy_async();
      mbar_arrive(&y_ready[tb]);
    };
    prepare(0);
    for (uint32_t t = 0; t < my_tiles; ++t) {
      FFD_STAMP(warp == 8 && t == 2, 400);
      if (t >= 1) finish(t - 1);
      FFD_STAMP(warp == 8 && t == 2, 401);
      if (t + 1 < my_tiles) prepare(t + 1);
      FFD_STAMP(warp == 8 && t == 2, 402);
    }
    finish(my_tiles - 1);
    if (lead) mbar_wait_relaxed(obf_empty, (my_tiles - 1) & 1, 200);  // drain
  } else {
    // ================= softmax groups: group wg owns quarter wg of the key columns of every virtual head =============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
    const int wg = (warp - 12) >> 2;       // 0..3: quarter of the key columns
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t scol0 = tmem_base + lane_base + COL_S + 2 * wg * NQ;  // this group's two score buffers; P goes over S
    // the four warps of a sub-partition (one per group) do identical work: start them a quarter of a head apart
#if !FFD_ATTN_RING
    __nanosleep(wg * FFD_STAGGER_NS);
#endif
    for (uint32_t n = 0; n < n_heads; ++n) {
      const uint32_t bf = n & 1;
      const uint32_t scol = scol0 + bf * NQ;
      FFD_STAMP(quad == 0 && n - 2u * NV < (uint32_t)NV, 512 + (wg * 32 + (n - 2u * NV)) * 4 + 0);
      mbar_wait(&s_full[2 * wg + bf], (n >> 1) & 1);
      tc_fence_after();
      FFD_STAMP(quad == 0 && n - 2u * NV < (uint32_t)NV, 512 + (wg * 32 + (n - 2u * NV)) * 4 + 1);
      uint32_t r[16 * KQ];
#pragma unroll
      for (int j = 0; j < KQ; ++j) FFD_TMEM_LD16(scol + 16 * j, (r + 16 * j));
      tmem_ld_wait();
      // row maximum over this quarter (independent chains).  Padded keys need no mask: slot 15 of every Q tile is the
      // constant 1 and slot 15 of a padded K row is -29952 (weight 0 in the combine).
      float mj[KQ];
#pragma unroll
      for (int j = 0; j < KQ; ++j) {
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < 16; e += 2) m = max3(m, __uint_as_float(r[16 * j + e]), __uint_as_float(r[16 * j + e + 1]));
        mj[j] = m;
      }
      float mb = mj[0];
#pragma unroll
      for (int j = 1; j < KQ; ++j) mb = fmaxf(mb, mj[j]);
      float* mslot = &mq[((n & 3) * 4 + wg) * TILE_M + row];
      *mslot = mb;
      float mc = -mb;
#if FFD_ATTN_RING
      // the exponential phase goes round the four groups: every sub-partition then has exactly one of its four warps on
      // the MUFU pipe while the other three load, reduce and store (left alone the four fall into step and the pipe
      // idles while all of them wait on TMEM round trips)
#if FFD_ATTN_RING == 2
      if (n | (wg >> 1)) mc = -named_bar_sync_reload(8 + (wg >> 1), 512, mslot);  // two sets of two groups take turns
#else
      if (n | wg) mc = -named_bar_sync_reload(8 + wg, 256, mslot);
#endif
#endif
      FFD_STAMP(quad == 0 && n - 2u * NV < (uint32_t)NV, 512 + (wg * 32 + (n - 2u * NV)) * 4 + 2);
      // p = exp2(s - m_q) as fp16 pairs, written over this warp's own score columns: 16 KQ fp32 -> 8 KQ packed columns
#pragma unroll
      for (int j = 0; j < KQ; ++j) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2)
          pk[e >> 1] = ((POLY_PAIRS >> (e >> 1)) & 1)
                           ? ex2_pack_f16_shift_poly(__uint_as_float(r[16 * j + e]), __uint_as_float(r[16 * j + e + 1]), mc)
                           : ex2_pack_f16_shift(__uint_as_float(r[16 * j + e]), __uint_as_float(r[16 * j + e + 1]), mc);
        FFD_TMEM_ST8(scol + 8 * j, pk);
      }
#if FFD_ATTN_RING
#if FFD_ATTN_RING == 2
      if (wg < 2 || n + 1 < n_heads) named_bar_arrive(8 + ((wg >> 1) ^ 1), 512);
#else
      if (wg != 3 || n + 1 < n_heads) named_bar_arrive(8 + ((wg + 1) & 3), 256);
#endif
#endif
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[2 * wg + bf]);
      FFD_STAMP(quad == 0 && n - 2u * NV < (uint32_t)NV, 512 + (wg * 32 + (n - 2u * NV)) * 4 + 3);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
#undef FFD_STAMP
}
#endif  // FFD_BUILD_ATTN_QUARTERS

// ---- packing ---------------------------------------------------------------------------------------
// WqP [QN n x DP k]: row 16 qt + g hd + c = Wq[(qt G + g) hd + c, :] * log2(e)/sqrt(hd), bias hi/lo in columns d,
// d+1; row 16 qt + 15 picks the constant-1 column of X.  WoP [DP n x DP k]: row n = Wo[n, :], bias hi/lo in
// columns d, d+1.   bf16 core-matrix order.
__global__ void pack_attn_kernel(const float* __restrict__ in_w, const float* __restrict__ in_b,
                                 const float* __restrict__ out_w, const float* __restrict__ out_b,
                                 op16* __restrict__ wq, op16* __restrict__ wo, int d, int H, int DP,
                                 int G, int QN) {
  const int hd = d / H;
  const float scale = 1.4426950408889634f / sqrtf((float)hd);
  const int total = QN * DP + DP * DP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool is_o = i >= QN * DP;
    const int e = is_o ? i - QN * DP : i;
    const int rows = is_o ? DP : QN;
    const int kg = e / (rows * 8), rem = e % (rows * 8);
    const int n = (rem / 64) * 8 + (rem % 64) / 8, k = kg * 8 + rem % 8;
    float v = 0.f;
    if (!is_o) {
      const int qt = n / HP, slot = n % HP;
      const int g = slot / hd, c = slot % hd, h = qt * G + g;
      if (slot == HP - 1) {
        v = (k == d) ? 1.f : 0.f;  // constant-1 slot of every Q tile (pairs with the key-padding slot of K)
      } else if (g < G && h < H) {
        const int r = h * hd + c;
        const float b = in_b[r] * scale;
        if (k < d) v = in_w[r * d + k] * scale;
        else if (k == d) v = round_op(b);
        else if (k == d + 1) v = b - round_op(b);
      }
    } else if (n < d) {
      const float b = out_b[n];
      if (k < d) v = out_w[n * d + k];
      else if (k == d) v = round_op(b);
      else if (k == d + 1) v = b - round_op(b);
    }
    (is_o ? wo : wq)[e] = to_op(v);
  }
}

// Weight images of the fused K/V projection (stock attention, hd <= 7), bf16 core-matrix order, bias hi/lo in
// columns d, d+1:
//   WkP [QN n x DP k]: row 16 qt + s = Wk[(qt G + s / hd) hd + s % hd, :] for s < G hd: the K rows of the G heads
//       of Q tile qt, already in the slot order of their K images (zero rows elsewhere);
//   WvP [8 H n x DP k]:  row 8 h + c = Wv[h hd + c, :] (c < hd); row 8 h + 7 picks the constant-1 column of X (the
//       row-sum slot of the V image).
__global__ void pack_kvproj_kernel(const float* __restrict__ in_w, const float* __restrict__ in_b,
                                   op16* __restrict__ wk, op16* __restrict__ wv, int d, int H,
                                   int DP, int G, int QN) {
  const int hd = d / H, NK = QN, NVR = 8 * H;
  const int total = (NK + NVR) * DP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const bool is_v = i >= NK * DP;
    const int e = is_v ? i - NK * DP : i;
    const int rows = is_v ? NVR : NK;
    const int kg = e / (rows * 8), rem = e % (rows * 8);
    const int n = (rem / 64) * 8 + (rem % 64) / 8, k = kg * 8 + rem % 8;
    float v = 0.f;
    int r = -1;  // row of in_proj (K rows at d.., V rows at 2d..)
    if (!is_v) {
      const int qt = n / 16, slot = n % 16, h = qt * G + slot / hd;
      if (slot < G * hd && h < H) r = d + h * hd + slot % hd;
    } else {
      const int h = n / 8, c = n % 8;
      if (c < hd) r = 2 * d + h * hd + c;
      else if (c == 7) v = (k == d) ? 1.f : 0.f;
    }
    if (r >= 0) {
      const float b = in_b[r];
      if (k < d) v = in_w[(int64_t)r * d + k];
      else if (k == d) v = round_op(b);
      else if (k == d + 1) v = b - round_op(b);
    }
    (is_v ? wv : wk)[e] = to_op(v);
  }
}

// Fused K/V projection + operand-image build for stock attention (hd <= 7): one CTA per 128-row tile of a sample.
//   K slots[128 x QN] = X[128 x DP] . WkP^T,  V dims[128 x 8 H] = X . WvP^T     (tcgen05, fp32 accumulate in TMEM)
// and every thread (= key row) writes its rows of the per-sample images (pack_kv_own_kernel's layout): two 16-byte
// K units per head from the accumulator (the MMA already produced the slot order; each head keeps its own slots,
// its tile mates' are zeroed), the V units after an 8 x 8 transpose through shared memory (a V unit holds 8
// consecutive keys of one dim).
template <int DP>
__global__ void __launch_bounds__(128, 4)
kv_own_tc_kernel(const float* __restrict__ x, const uint8_t* __restrict__ wkimg, uint8_t* __restrict__ img, int d,
                 int H, int hd, int G, int QN, int L, int NCH, int CK, int tps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int NK = QN, NVR = 8 * H;
  const int wk_bytes = NK * DP * 2, wv_bytes = NVR * DP * 2;
  uint8_t* xo = smem;                                   // bf16 A operand [128 x DP]
  uint8_t* wk = xo + TILE_M * DP * 2;                   // WkP, then WvP
  uint8_t* stage = wk + wk_bytes + wv_bytes;            // [4 warps][32 keys][8 dims] fp16
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 4 * 32 * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  uint64_t* w_full = bars;
  uint64_t* k_done = bars + 1;
  uint64_t* v_done = bars + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  const int64_t sample = blockIdx.x / tps;
  const int part = (int)(blockIdx.x - sample * tps);
  const int rows = L - part * TILE_M < TILE_M ? L - part * TILE_M : TILE_M;
  const bool valid = row < rows;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(k_done, 1);
    mbar_init(v_done, 1);
    fence_barrier_init();
    mbar_expect_tx(w_full, wk_bytes + wv_bytes);
    bulk_load(wk, wkimg, wk_bytes + wv_bytes, w_full);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);  // K slots, then V dims: four CTAs per SM
  // stage this tile's rows as the bf16 A operand; rows beyond the sample are all zero (their constant columns too,
  // so that their V row-sum slot stays 0)
  if ((d & 7) == 0) {
    // A warp's 32 rows are one contiguous block of global memory: lane l reads its 16-byte chunks l, l + 32, ...
    // (coalesced); two neighbouring chunks of a row make one 16-byte bf16 unit of the operand (even lanes pack).
    const int n4 = d >> 2;
    const int wrows = rows - warp * 32 < 32 ? rows - warp * 32 : 32;  // valid rows of this warp (may be <= 0)
    const float4* src = reinterpret_cast<const float4*>(x + (sample * L + (int64_t)part * TILE_M + warp * 32) * d);
    const int step_r = 32 / n4, step_c = 32 % n4;
    int r = lane / n4, c4 = lane % n4;
#pragma unroll 6
    for (int q = lane; q < 32 * n4; q += 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < wrows) a = src[q];
      float4 b4;
      b4.x = __shfl_down_sync(0xffffffffu, a.x, 1);
      b4.y = __shfl_down_sync(0xffffffffu, a.y, 1);
      b4.z = __shfl_down_sync(0xffffffffu, a.z, 1);
      b4.w = __shfl_down_sync(0xffffffffu, a.w, 1);
      if ((lane & 1) == 0) {
        const int trow = warp * 32 + r;
        uint4 pk;
        pk.x = pack_op(a.x, a.y);
        pk.y = pack_op(a.z, a.w);
        pk.z = pack_op(b4.x, b4.y);
        pk.w = pack_op(b4.z, b4.w);
        *reinterpret_cast<uint4*>(xo + (uint32_t)(trow >> 3) * 128 + (uint32_t)(trow & 7) * 16 +
                                  (c4 >> 1) * (TILE_M * 16)) = pk;
      }
      r += step_r, c4 += step_c;
      if (c4 >= n4) c4 -= n4, ++r;
    }
    // constant columns (d, d + 1 = 1 for valid rows) and the zero padding up to DP
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    for (int kg = d >> 3; kg < DP / 8; ++kg) {
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      if (kg == (d >> 3) && valid) pk.x = pack_op(1.f, 1.f);
      *reinterpret_cast<uint4*>(xo + row_off + kg * (TILE_M * 16)) = pk;
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(x + (sample * L + (int64_t)part * TILE_M + row) * d);
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    constexpr int NG = DP / 8;
#pragma unroll 5
    for (int kg = 0; kg < NG; ++kg) {
      const int c = kg * 8;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
      if (valid) {
        if (c < d) a = src[2 * kg];
        else if (c == d) a.x = 1.f, a.y = 1.f;
        if (c + 4 < d) b4 = src[2 * kg + 1];
        else if (c + 4 == d) b4.x = 1.f, b4.y = 1.f;
      }
      uint4 q;
      q.x = pack_op(a.x, a.y);
      q.y = pack_op(a.z, a.w);
      q.z = pack_op(b4.x, b4.y);
      q.w = pack_op(b4.z, b4.w);
      *reinterpret_cast<uint4*>(xo + row_off + kg * (TILE_M * 16)) = q;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t LBO128 = 128 * 16;
  if (warp == 0) {
    mbar_wait(w_full, 0);
    tc_fence_after();
    umma_ss_chain<DP / 16>(tmem_base, make_desc(smem_u32(xo), LBO128, 128), (2 * LBO128) >> 4,
                           make_desc(smem_u32(wk), NK * 16, 128), (2 * NK * 16) >> 4, make_idesc(TILE_M, NK), 0);
    umma_commit(k_done);
  }
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const size_t k_bytes = (size_t)CK * HP * 2, img_stride = 2 * k_bytes;
  uint8_t* simg = img + (size_t)sample * H * NCH * img_stride;
  // keys this thread writes: its own row, plus (last tile of the sample) the padded keys beyond the tiles
  const int key = part * TILE_M + row;
  const int n_keys = NCH * CK;
  // ---- K images ----
  mbar_wait(k_done, 0);
  tc_fence_after();
  for (int qt = 0; qt * 16 < NK; ++qt) {
    uint32_t r[16];
    FFD_TMEM_LD16(tmem_base + lane_base + 16 * qt, r);
    tmem_ld_wait();
    if (key < n_keys) {
      const int ch = key / CK, kl = key - ch * CK;
      for (int g = 0; g < G; ++g) {
        const int h = qt * G + g;
        if (h >= H) break;
        const int lo = g * hd, hi = lo + hd;  // this head's slots
        float v[16];
#pragma unroll
        for (int sl = 0; sl < 15; ++sl) v[sl] = (sl >= lo && sl < hi) ? __uint_as_float(r[sl]) : 0.f;
        v[15] = valid ? 0.f : -29952.f;  // padded keys: score -29952 -> weight 0
        uint4 q0, q1;
        q0.x = pack_op(v[0], v[1]), q0.y = pack_op(v[2], v[3]), q0.z = pack_op(v[4], v[5]);
        q0.w = pack_op(v[6], v[7]);
        q1.x = pack_op(v[8], v[9]), q1.y = pack_op(v[10], v[11]), q1.z = pack_op(v[12], v[13]);
        q1.w = pack_op(v[14], v[15]);
        uint8_t* kimg = simg + (size_t)(h * NCH + ch) * img_stride;
        *reinterpret_cast<uint4*>(kimg + (size_t)kl * 16) = q0;
        *reinterpret_cast<uint4*>(kimg + (size_t)(CK + kl) * 16) = q1;
      }
    }
  }
  if (part == tps - 1) {
    // padded keys not covered by any tile (only when NCH CK > 128 tps)
    const uint4 z0 = make_uint4(0u, 0u, 0u, 0u);
    uint4 z1 = z0;
    z1.w = pack_op(0.f, -29952.f);
    for (int k2 = tps * TILE_M + row; k2 < n_keys; k2 += TILE_M) {
      const int ch = k2 / CK, kl = k2 - ch * CK;
      for (int h = 0; h < H; ++h) {
        uint8_t* kimg = simg + (size_t)(h * NCH + ch) * img_stride;
        *reinterpret_cast<uint4*>(kimg + (size_t)kl * 16) = z0;
        *reinterpret_cast<uint4*>(kimg + (size_t)(CK + kl) * 16) = z1;
        if ((kl & 7) == 0) {  // V units of this group of 8 padded keys: 16 dims of zeros
          uint8_t* vimg = kimg + k_bytes + (size_t)(kl >> 3) * 256;
          for (int u = 0; u < 16; ++u) *reinterpret_cast<uint4*>(vimg + u * 16) = z0;
        }
      }
    }
  }
  // ---- V images: 8 x 8 transpose of (key, dim) through shared memory, per group of 8 keys ----
  tc_fence_before();
  __syncthreads();  // every thread has read its K accumulators: the columns are reused for V
  if (warp == 0) {
    tc_fence_after();
    umma_ss_chain<DP / 16>(tmem_base, make_desc(smem_u32(xo), LBO128, 128), (2 * LBO128) >> 4,
                           make_desc(smem_u32(wk + wk_bytes), NVR * 16, 128), (2 * NVR * 16) >> 4,
                           make_idesc(TILE_M, NVR), 0);
    umma_commit(v_done);
  }
  mbar_wait(v_done, 0);
  tc_fence_after();
  uint4* my_stage = reinterpret_cast<uint4*>(stage + warp * 512);
  const uint16_t* col = reinterpret_cast<const uint16_t*>(stage + warp * 512) + (lane & ~7) * 8 + (lane & 7);
  for (int h0 = 0; h0 < H; h0 += 2) {
    // two heads per TMEM load: columns 8 h0 .. 8 h0 + 15
    uint32_t r[16];
    FFD_TMEM_LD16(tmem_base + lane_base + 8 * h0, r);
    tmem_ld_wait();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = h0 + hh;
      if (h < H) {
        uint4 q;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(__uint_as_float(r[8 * hh + 1])), "f"(__uint_as_float(r[8 * hh + 0])));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(__uint_as_float(r[8 * hh + 3])), "f"(__uint_as_float(r[8 * hh + 2])));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.z) : "f"(__uint_as_float(r[8 * hh + 5])), "f"(__uint_as_float(r[8 * hh + 4])));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.w) : "f"(__uint_as_float(r[8 * hh + 7])), "f"(__uint_as_float(r[8 * hh + 6])));
        __syncwarp();
        my_stage[lane] = q;  // [key][8 dims]
        __syncwarp();
        // lane 8 i + j gathers dim j of keys 8 i .. 8 i + 7
        uint4 t;
        t.x = (uint32_t)col[0] | ((uint32_t)col[8] << 16);
        t.y = (uint32_t)col[16] | ((uint32_t)col[24] << 16);
        t.z = (uint32_t)col[32] | ((uint32_t)col[40] << 16);
        t.w = (uint32_t)col[48] | ((uint32_t)col[56] << 16);
        const int key0 = part * TILE_M + (row & ~7);  // first key of this lane's group
        if (key0 < n_keys) {
          const int ch = key0 / CK, kl = key0 - ch * CK;
          uint8_t* vimg = simg + (size_t)(h * NCH + ch) * img_stride + k_bytes + (size_t)(kl >> 3) * 256;
          *reinterpret_cast<uint4*>(vimg + (lane & 7) * 16) = t;
          *reinterpret_cast<uint4*>(vimg + (8 + (lane & 7)) * 16) = make_uint4(0u, 0u, 0u, 0u);  // dims 8..15
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// Images per (layer, head h, key chunk c), keys [c CK, (c+1) CK):
//   K: B operand [CK n(key) x 16 k(slot)], bf16: head h's dims in the slots of its place g = h % G inside its Q
//      tile, zeros in the slots of its tile mates; slot 15 of a padded key row is -29952 (slot 15 of Q is 1).
//   V: B operand [16 n(dim) x CK k(key)], fp16 (P is fp16), with row 7 (hd <= 7) or row 15 set to 1 for the valid
//      keys so that this column of O = P . V^T is the softmax row sum.
__global__ void pack_kv_kernel(const float* __restrict__ kc, const float* __restrict__ vc,
                               uint16_t* __restrict__ img, int n_lh, int L, int hd, int H, int G, int NCH, int CK) {
  const int per = CK * HP;  // elements per image
  const int64_t total = (int64_t)n_lh * NCH * 2 * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int lhc = (int)(i / (2 * per));
    const int lh = lhc / NCH, ch = lhc % NCH;
    const int g = (lh % H) % G;
    const int r = (int)(i % (2 * per));
    const bool is_v = r >= per;
    const int e = is_v ? r - per : r;
    const int rows = is_v ? HP : CK;
    const int kg = e / (rows * 8), rem = e % (rows * 8);
    const int n = (rem / 64) * 8 + (rem % 64) / 8, k = kg * 8 + rem % 8;
    const int key = ch * CK + (is_v ? k : n), slot = is_v ? n : k;
    float v = 0.f;
    if (is_v) {
      if (key < L && slot < hd) v = vc[((int64_t)lh * L + key) * hd + slot];
      if (key < L && slot == (hd <= 7 ? 7 : HP - 1)) v = 1.f;  // that column of O = softmax row sum
    } else {
      const int dim = slot - g * hd;
      if (key < L && slot < HP - 1 && dim >= 0 && dim < hd) v = kc[((int64_t)lh * L + key) * hd + dim];
      if (key >= L && slot == HP - 1) v = -29952.f;  // padded keys: score = -29952 -> weight 0
    }
    img[i] = is_v ? __half_as_ushort(__float2half_rn(v)) : op_bits(v);
  }
}

// Mixed branch (cached_transformer.py:259-305): a sample's own K/V rows only for the tokens of the recompute set S
// (prefix [0, n_prefix) or mask); every other key row of the per-sample images is replaced by the row of the shared
// sample-0 images of this layer.  One thread per 16-byte core-matrix row: a K-image row belongs to one key, a
// V-image row holds 8 consecutive keys of one dim.
__global__ void patch_kv_own_kernel(uint4* __restrict__ own, const uint4* __restrict__ shared_layer, int64_t B, int NV,
                                    int NCH, int CK, int L, int n_prefix, const int32_t* __restrict__ mask) {
  const int units = 4 * CK;
  const int64_t total = B * NV * units;
  auto in_set = [&](int key) { return key < L && (mask ? mask[key] != 0 : key < n_prefix); };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % units);
    const int hc = (int)((i / units) % NV);
    const int ch = hc % NCH;
    const uint4 sh = shared_layer[(int64_t)hc * units + u];
    if (u < 2 * CK) {
      if (!in_set(ch * CK + u % CK)) own[i] = sh;
    } else {
      const int key0 = ch * CK + ((u - 2 * CK) / 16) * 8;
      uint4 o = own[i];
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
      const uint32_t* sw = reinterpret_cast<const uint32_t*>(&sh);
      bool changed = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool k0 = in_set(key0 + 2 * j), k1 = in_set(key0 + 2 * j + 1);
        const uint32_t v = (k0 ? ow[j] : sw[j]) & 0xFFFFu, w = (k1 ? ow[j] : sw[j]) & 0xFFFF0000u;
        changed |= !k0 || !k1;
        ow[j] = v | w;
      }
      if (changed) own[i] = o;
    }
  }
}

// The same images for every sample's OWN keys (stock attention): K/V rows come from the projected tokens
// src[(b L + key) * ld + k_off / v_off + h hd + dim].  One thread writes one 16-byte core-matrix row.
__global__ void pack_kv_own_kernel(const float* __restrict__ src, int ld, int k_off, int v_off,
                                   uint4* __restrict__ img, int64_t B, int L, int hd, int H, int G, int NCH, int CK) {
  const int units = 4 * CK;  // per (sample, head, chunk): 2 CK of the K image, then 2 CK of the V image
  const int64_t total = B * H * NCH * units;
  const int rs = hd <= 7 ? 7 : HP - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % units);
    const int64_t bhc = i / units;
    const int ch = (int)(bhc % NCH);
    const int h = (int)((bhc / NCH) % H);
    const int64_t b = bhc / ((int64_t)NCH * H);
    const int g = h % G;
    float v[8];
    if (u < 2 * CK) {
      const int kg = u / CK, key = ch * CK + u % CK;
      const float* row = src + (b * L + key) * ld + k_off + h * hd;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int slot = kg * 8 + j, dim = slot - g * hd;
        v[j] = 0.f;
        if (key < L && slot < HP - 1 && dim >= 0 && dim < hd) v[j] = row[dim];
        if (key >= L && slot == HP - 1) v[j] = -29952.f;
      }
      uint4 q;
      q.x = pack_op(v[0], v[1]), q.y = pack_op(v[2], v[3]), q.z = pack_op(v[4], v[5]), q.w = pack_op(v[6], v[7]);
      img[i] = q;
    } else {
      const int uu = u - 2 * CK;
      const int kg = uu / HP, dim = uu % HP;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = ch * CK + kg * 8 + j;
        v[j] = 0.f;
        if (key < L) v[j] = dim < hd ? src[(b * L + key) * ld + v_off + h * hd + dim] : (dim == rs ? 1.f : 0.f);
      }
      uint4 q;
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(v[1]), "f"(v[0]));
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(v[3]), "f"(v[2]));
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.z) : "f"(v[5]), "f"(v[4]));
      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.w) : "f"(v[7]), "f"(v[6]));
      img[i] = q;
    }
  }
}

}  // namespace tc

// Geometry of the attention kernel for a model: G heads per Q tile, QN Q columns, NCH key chunks of CK keys.
struct AttnGeom {
  int hd, G, QN, NCH, CK, KQ, NV;
};

static AttnGeom attn_geom(const ffd_dims& m) {
  AttnGeom g;
  g.hd = m.d_model / m.n_head;
  g.G = g.hd <= 15 ? 15 / g.hd : 1;
  if (g.G > m.n_head) g.G = m.n_head;
  g.QN = (m.n_head + g.G - 1) / g.G * tc::HP;
  g.NCH = attn_key_chunks(m.max_len);
  g.CK = attn_chunk_keys(m.max_len);
  g.KQ = g.CK / 64;
  g.NV = m.n_head * g.NCH;
  return g;
}

bool tc_attn_supported(const ffd_dims& m) {
  if (!tc_supported(m)) return false;
  const AttnGeom g = attn_geom(m);
  if (g.hd > 15 || g.QN > 256 || g.NV < 2) return false;
  return tc::attn_smem(tc::dp_of(m.d_model), g.QN, g.CK, m.d_model).total <= 232448;
}

// fused K/V projection images (WkP, WvP) exist for hd <= 7 and at most 16 heads (TMEM plan of kv_own_tc_kernel)
static bool kvproj_fused(const ffd_dims& m) {
  const AttnGeom g = attn_geom(m);
  return g.hd <= 7 && g.QN <= 128 && 8 * m.n_head <= 128 && m.n_head % 2 == 0;
}

int64_t tc_attn_image_bytes(const ffd_dims& m) {
  const int DP = tc::dp_of(m.d_model), QN = attn_geom(m).QN;
  return (int64_t)(QN * DP + DP * DP + (kvproj_fused(m) ? (QN + 8 * m.n_head) * DP : 0)) * 2;
}

int64_t tc_kv_image_bytes(const ffd_dims& m) {
  const AttnGeom g = attn_geom(m);
  return (int64_t)m.n_layers * g.NV * 2 * g.CK * tc::HP * 2;
}

int tc_pack_attn(const ffd_dims& m, const float* weights, void* img_base, int64_t layer_stride_bytes,
                 int64_t attn_off_bytes, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const AttnGeom g = attn_geom(m);
  const int DP = tc::dp_of(m.d_model);
  for (int l = 0; l < m.n_layers; ++l) {
    const float* w = weights + ar.layer0 + (int64_t)l * ar.layer_stride;
    tc::op16* wq = (tc::op16*)((uint8_t*)img_base + l * layer_stride_bytes + attn_off_bytes);
    tc::pack_attn_kernel<<<32, 256, 0, s>>>(w + ar.in_w, w + ar.in_b, w + ar.out_w, w + ar.out_b, wq,
                                            wq + (size_t)g.QN * DP, m.d_model, m.n_head, DP, g.G, g.QN);
    FFD_CHECK_LAUNCH("pack_attn_kernel");
    if (kvproj_fused(m)) {
      tc::op16* wk = wq + (size_t)g.QN * DP + (size_t)DP * DP;
      tc::pack_kvproj_kernel<<<32, 256, 0, s>>>(w + ar.in_w, w + ar.in_b, wk, wk + (size_t)g.QN * DP, m.d_model,
                                                m.n_head, DP, g.G, g.QN);
      FFD_CHECK_LAUNCH("pack_kvproj_kernel");
    }
  }
  return FFD_OK;
}

int tc_pack_kv(const ffd_dims& m, const float* kc, const float* vc, void* kvimg, cudaStream_t s) {
  const AttnGeom g = attn_geom(m);
  tc::pack_kv_kernel<<<148, 256, 0, s>>>(kc, vc, (uint16_t*)kvimg, m.n_layers * m.n_head, m.max_len, g.hd, m.n_head,
                                         g.G, g.NCH, g.CK);
  FFD_CHECK_LAUNCH("pack_kv_kernel");
  return FFD_OK;
}

int tc_pack_kv_layer(const ffd_dims& m, int layer, const float* kc, const float* vc, void* kvimg, cudaStream_t s) {
  const AttnGeom g = attn_geom(m);
  const int64_t cache_off = (int64_t)layer * m.n_head * m.max_len * g.hd;
  const int64_t img_off = (int64_t)layer * m.n_head * g.NCH * 2 * g.CK * tc::HP;  // 16-bit elements
  tc::pack_kv_kernel<<<16, 256, 0, s>>>(kc + cache_off, vc + cache_off, (uint16_t*)kvimg + img_off, m.n_head,
                                        m.max_len, g.hd, m.n_head, g.G, g.NCH, g.CK);
  FFD_CHECK_LAUNCH("pack_kv_kernel");
  return FFD_OK;
}

// the same images in chunks of 96 keys (virtual heads of the fused layer kernel, ffd_tc_layer.cu)
int tc_pack_kv96(const ffd_dims& m, const float* kc, const float* vc, void* kvimg96, cudaStream_t s) {
  const AttnGeom g = attn_geom(m);
  const int NCH = (m.max_len + 95) / 96;
  tc::pack_kv_kernel<<<148, 256, 0, s>>>(kc, vc, (uint16_t*)kvimg96, m.n_layers * m.n_head, m.max_len, g.hd, m.n_head,
                                         g.G, NCH, 96);
  FFD_CHECK_LAUNCH("pack_kv_kernel");
  return FFD_OK;
}

int tc_pack_kv96_layer(const ffd_dims& m, int layer, const float* kc, const float* vc, void* kvimg96, cudaStream_t s) {
  const AttnGeom g = attn_geom(m);
  const int NCH = (m.max_len + 95) / 96;
  const int64_t cache_off = (int64_t)layer * m.n_head * m.max_len * g.hd;
  const int64_t img_off = (int64_t)layer * m.n_head * NCH * 2 * 96 * tc::HP;  // 16-bit elements
  tc::pack_kv_kernel<<<16, 256, 0, s>>>(kc + cache_off, vc + cache_off, (uint16_t*)kvimg96 + img_off, m.n_head,
                                        m.max_len, g.hd, m.n_head, g.G, NCH, 96);
  FFD_CHECK_LAUNCH("pack_kv_kernel");
  return FFD_OK;
}

long long* g_attn_dbg = nullptr;  // optional device buffer of clock64 stamps (profiling aid)

// attention kernel variant: 1 = v12 (four independent quarter streams), 0 = v11 (lock-step heads); -1 = FFD_ATTN_QUARTERS
// or the default
#ifndef FFD_ATTN_QUARTERS_DEFAULT
#define FFD_ATTN_QUARTERS_DEFAULT 0
#endif
static int g_attn_quarters_mode = -1;
void tc_debug_set_attn_quarters(int mode) { g_attn_quarters_mode = mode; }
static bool attn_quarters_enabled() {
  int mode = g_attn_quarters_mode;
  if (mode < 0) {
    static const int env_mode = [] {
      const char* e = getenv("FFD_ATTN_QUARTERS");
      return e ? (e[0] == '1' ? 1 : 0) : FFD_ATTN_QUARTERS_DEFAULT;
    }();
    mode = env_mode;
  }
  return mode == 1;
}

template <int DP, int KQ>
static int launch_attn(const ffd_dims& m, const AttnGeom& g, const uint8_t* wqimg, const uint8_t* kvimg_layer,
                       const float* lnw, const float* lnb, const float* x, float* y, int64_t M, int own,
                       cudaStream_t s) {
  static int n_sm = 0;
  const tc::AttnSmem sm = tc::attn_smem(DP, g.QN, g.CK, m.d_model);
  if (n_sm == 0) {
    int dev = 0;
    FFD_CHECK_CUDA(cudaGetDevice(&dev));
    FFD_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::attn_cached_tc_kernel<DP, KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      sm.total));
  // own = 1: per-sample K/V images, tiles aligned to samples
  const int tps = (m.max_len + tc::TILE_M - 1) / tc::TILE_M;
  const int64_t n_tiles = own ? M / m.max_len * tps : (M + tc::TILE_M - 1) / tc::TILE_M;
  const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
  if (attn_quarters_enabled()) {
#ifdef FFD_BUILD_ATTN_QUARTERS
    // v12: four independent quarter streams (same arguments, identical arithmetic per element)
    FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::attn_quarters_tc_kernel<DP, KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        sm.total));
    tc::attn_quarters_tc_kernel<DP, KQ><<<grid, tc::ATT2_THREADS, sm.total, s>>>(
        x, y, wqimg, wqimg + (size_t)g.QN * DP * 2, kvimg_layer, lnw, lnb, M, m.d_model, m.n_head, g.hd, g.G, g.QN,
        g.NCH, n_tiles, 1e-5f, own ? m.max_len : 0, tps, g_attn_dbg);
    FFD_CHECK_LAUNCH("attn_quarters_tc_kernel");
    return FFD_OK;
#else
    set_error("the quarter-stream attention kernel is an experiment: rebuild with -DFFD_BUILD_ATTN_QUARTERS");
    return FFD_ERR_ARG;
#endif
  }
  tc::attn_cached_tc_kernel<DP, KQ><<<grid, tc::ATT_THREADS, sm.total, s>>>(
      x, y, wqimg, wqimg + (size_t)g.QN * DP * 2, kvimg_layer, lnw, lnb, M, m.d_model, m.n_head, g.hd, g.G, g.QN,
      g.NCH, n_tiles, 1e-5f, own ? m.max_len : 0, tps, g_attn_dbg);
  FFD_CHECK_LAUNCH("attn_cached_tc_kernel");
  return FFD_OK;
}

template <int DP>
static int launch_attn_kq(const ffd_dims& m, const AttnGeom& g, const uint8_t* wqimg, const uint8_t* kvimg_layer,
                          const float* lnw, const float* lnb, const float* x, float* y, int64_t M, int own,
                          cudaStream_t s) {
  switch (g.KQ) {
    case 1: return launch_attn<DP, 1>(m, g, wqimg, kvimg_layer, lnw, lnb, x, y, M, own, s);
    case 2: return launch_attn<DP, 2>(m, g, wqimg, kvimg_layer, lnw, lnb, x, y, M, own, s);
    case 3: return launch_attn<DP, 3>(m, g, wqimg, kvimg_layer, lnw, lnb, x, y, M, own, s);
  }
  set_error("tensor-core attention: unsupported key chunk %d", g.CK);
  return FFD_ERR_UNSUPPORTED;
}

static int attn_dispatch(const ffd_dims& m, const AttnGeom& g, const uint8_t* wq, const uint8_t* kvl, const float* w,
                         const Arena& ar, const float* x, float* y, int64_t M, int own, cudaStream_t s) {
  switch (tc::dp_of(m.d_model)) {
    case 16: return launch_attn_kq<16>(m, g, wq, kvl, w + ar.n1w, w + ar.n1b, x, y, M, own, s);
    case 32: return launch_attn_kq<32>(m, g, wq, kvl, w + ar.n1w, w + ar.n1b, x, y, M, own, s);
    case 48: return launch_attn_kq<48>(m, g, wq, kvl, w + ar.n1w, w + ar.n1b, x, y, M, own, s);
    case 64: return launch_attn_kq<64>(m, g, wq, kvl, w + ar.n1w, w + ar.n1b, x, y, M, own, s);
    case 80: return launch_attn_kq<80>(m, g, wq, kvl, w + ar.n1w, w + ar.n1b, x, y, M, own, s);
  }
  set_error("tensor-core attention: unsupported d_model %d", m.d_model);
  return FFD_ERR_UNSUPPORTED;
}

int tc_attn_cached(const ffd_dims& m, const void* attn_img, const void* kvimg, const float* weights, int layer,
                   const float* x, float* y, int64_t M, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const AttnGeom g = attn_geom(m);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const uint8_t* kvl = (const uint8_t*)kvimg + (size_t)layer * g.NV * 2 * g.CK * tc::HP * 2;
  return attn_dispatch(m, g, (const uint8_t*)attn_img, kvl, w, ar, x, y, M, 0, s);
}

int64_t tc_kv_own_image_bytes(const ffd_dims& m, int64_t B) {
  const AttnGeom g = attn_geom(m);
  return B * g.NV * 2 * g.CK * tc::HP * 2;
}

template <int DP>
static int launch_kv_own(const ffd_dims& m, const AttnGeom& g, const uint8_t* wkimg, const float* x, void* img,
                         int64_t B, cudaStream_t s) {
  const int smem = tc::TILE_M * DP * 2 + (g.QN + 8 * m.n_head) * DP * 2 + 4 * 32 * 16 + 64;
  FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::kv_own_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int tps = (m.max_len + tc::TILE_M - 1) / tc::TILE_M;
  tc::kv_own_tc_kernel<DP><<<(unsigned)(B * tps), 128, smem, s>>>(x, wkimg, (uint8_t*)img, m.d_model, m.n_head, g.hd,
                                                                   g.G, g.QN, m.max_len, g.NCH, g.CK, tps);
  FFD_CHECK_LAUNCH("kv_own_tc_kernel");
  return FFD_OK;
}

bool tc_attn_own_fused(const ffd_dims& m) { return kvproj_fused(m); }

int tc_attn_own(const ffd_dims& m, const void* attn_img, void* kvimg_own, const float* kv_src, int ld, int k_off,
                int v_off, const float* weights, int layer, const float* x, float* y, int64_t M, cudaStream_t s,
                const void* kvimg_shared, int n_prefix, const int32_t* own_mask) {
  const Arena ar = arena_of(m);
  const AttnGeom g = attn_geom(m);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const int64_t B = M / m.max_len;
  if (kv_src == nullptr) {
    // K/V projection fused into the image build (tcgen05): images straight from x
    const int DP = tc::dp_of(m.d_model);
    const uint8_t* wk = (const uint8_t*)attn_img + (size_t)(g.QN * DP + DP * DP) * 2;
    int rc = FFD_ERR_UNSUPPORTED;
    switch (DP) {
      case 16: rc = launch_kv_own<16>(m, g, wk, x, kvimg_own, B, s); break;
      case 32: rc = launch_kv_own<32>(m, g, wk, x, kvimg_own, B, s); break;
      case 48: rc = launch_kv_own<48>(m, g, wk, x, kvimg_own, B, s); break;
      case 64: rc = launch_kv_own<64>(m, g, wk, x, kvimg_own, B, s); break;
      case 80: rc = launch_kv_own<80>(m, g, wk, x, kvimg_own, B, s); break;
    }
    if (rc) return rc;
  } else {
    tc::pack_kv_own_kernel<<<148 * 16, 256, 0, s>>>(kv_src, ld, k_off, v_off, (uint4*)kvimg_own, B, m.max_len, g.hd,
                                                     m.n_head, g.G, g.NCH, g.CK);
    FFD_CHECK_LAUNCH("pack_kv_own_kernel");
  }
  if (kvimg_shared) {
    const uint8_t* kvl = (const uint8_t*)kvimg_shared + (size_t)layer * g.NV * 2 * g.CK * tc::HP * 2;
    tc::patch_kv_own_kernel<<<148 * 16, 256, 0, s>>>((uint4*)kvimg_own, (const uint4*)kvl, B, g.NV, g.NCH, g.CK,
                                                      m.max_len, n_prefix, own_mask);
    FFD_CHECK_LAUNCH("patch_kv_own_kernel");
  }
  return attn_dispatch(m, g, (const uint8_t*)attn_img, (const uint8_t*)kvimg_own, w, ar, x, y, M, 1, s);
}

}  // namespace ffd
