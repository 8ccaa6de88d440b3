// tcgen05 / TMEM kernel for the attention half of a PURE-CACHE encoder layer (998 of 1000 default steps):
//   y = LN1(x + out_proj(softmax(q K_c^T / sqrt(hd)) V_c)),  q = x Wq^T + bq
//   cached_transformer.py:228-258 (pure-cache K/V), :309-321 (attention, out-proj, residual, norm1)
// K_c / V_c are the sample-0 cache rows shared by the whole batch (caching.py:326-328), so every token
// row is independent and the work is a chain of GEMMs over 128-token tiles:
//   Q[128 x 16H]   = X[128 x DP] . WqP^T          (heads padded to 16 columns; log2(e)/sqrt(hd) and bq folded in)
//   S_h[128 x NKV] = Q_h[128 x 16] . K_h^T        one MMA per head, fp32 scores in TMEM
//   P_h            = exp2(S_h - m)                softmax warps: TMEM -> registers -> TMEM (fp16 pairs, in place over
//                                                 S_h).  The key columns are split in four quarters, one per softmax
//                                                 group; each quarter uses ITS OWN row maximum m_q as the shift and
//                                                 gets its own O accumulator, and the quarters are combined when O is
//                                                 read out (flash-attention style split-K: exact, no exchange needed)
//   O_h[128 x 16]  = P_h . [V_h | 1]              A operand read from TMEM; the extra ones column of the V image
//                                                 makes the tensor core return the softmax row sum as well
//   OUT[128 x DP]  = concat_h(O_h / rowsum_h) . WoP^T   (bo folded in) -> + residual -> LayerNorm
// The fp32 token tile enters and leaves shared memory as one bulk async copy each way; the [tokens x heads x L]
// score tensor the reference materialises twice per layer (27.5 GB each at B=16384) lives in TMEM.
//
// Warp roles (22 warps): 0 = bulk-copy producer (Wq/Wo once, token tiles, K/V head images through rings),
// 1 = TMEM allocator + MMA issuer, 2..17 = four softmax groups of four warps, 18..21 = O read-out warps
// (combine the quarter accumulators, normalise, write the out-proj operand) so the softmax groups never wait.  All four groups work on the same
// head, each on one quarter of the key columns, with no synchronisation among them inside a head; the
// score buffers alternate between consecutive heads, so the tensor core computes S_{h+1} and P_{h-1} V_{h-1}
// while the softmax warps are busy with head h.  The softmax is bound by MUFU.EX2 (16 lanes/clk/SM).
#include <cuda_fp16.h>

#include "ffd_tc_ptx.cuh"

// Share of the softmax exponentials evaluated on the FMA pipe (bit mask over the 8 fp16 pairs of a 16-column
// chunk).  Measured on B200 at ECG shapes: 0x00 -> 2.96 ms/launch, 0x88 (25 %) -> 3.09, 0x92 -> 3.23, 0xAA (50 %)
// -> 3.47: FP32 FMA issues at half rate per sub-partition, so the polynomial costs more than the MUFU time it
// saves.  Kept as a build-time knob (tools/polytest.sh), off by default.
#ifndef FFD_POLY_PAIRS
#define FFD_POLY_PAIRS 0x00u
#endif

namespace ffd {
namespace tc {

constexpr int ATT_THREADS = 704;
constexpr int KV_STAGES = 3;   // K image ring
constexpr int V_STAGES = 2;    // V image ring (shared memory is full)
constexpr int HP = 16;  // padded head dim (UMMA K of the score MMA, N of the PV MMA)
// of the 8 fp16 pairs of a 16-column chunk, the pairs in this bit mask are exponentiated on the FMA pipe
constexpr uint32_t POLY_PAIRS = FFD_POLY_PAIRS;

// TMEM columns
constexpr uint32_t COL_S1 = 0;     // Q accumulators, then score buffer of the odd heads
constexpr uint32_t COL_S0 = 192;   // score buffer of the even heads
constexpr uint32_t COL_O = 384;    // O accumulators: buffer b, quarter q at COL_O + 64 b + 16 q
constexpr uint32_t COL_OUT = 192;  // out-proj accumulators [192, 192 + DP): aliases S0, free at that point

struct AttnSmem {
  int x32, xo, qbf, wq, wo, kring, vring, consts, pmax, bars, tmem, total;
  int tile_bytes, tile_stride, k_bytes, v_bytes, wq_bytes, wo_bytes;
};

__host__ __device__ inline AttnSmem attn_smem(int DP, int QN, int NKV, int d) {
  AttnSmem s;
  int o = 0;
  auto take = [&](int bytes) {
    int r = o;
    o += (bytes + 127) / 128 * 128;
    return r;
  };
  s.tile_bytes = TILE_M * d * 4;
  s.tile_stride = (s.tile_bytes + 127) / 128 * 128;
  s.k_bytes = NKV * HP * 2;
  s.v_bytes = HP * NKV * 2;
  s.wq_bytes = QN * DP * 2;
  s.wo_bytes = DP * DP * 2;
  s.x32 = take(2 * s.tile_stride);      // fp32 token tiles (ping-pong): input, residual, output staging
  s.xo = take(TILE_M * DP * 2);         // bf16 A operand of the Q projection (X of the current/next tile)
  s.qbf = take(TILE_M * QN * 2);        // bf16 Q head tiles; its first DP columns double as the out-proj A operand
  s.wq = take(s.wq_bytes);
  s.wo = take(s.wo_bytes);
  s.kring = take(KV_STAGES * s.k_bytes);
  s.vring = take(V_STAGES * s.v_bytes);
  s.consts = take(2 * DP * 4);
  s.pmax = take(4 * 4 * TILE_M * 4);    // quarter row maxima [head & 3][quarter][row]
  s.bars = take(64 * 8);
  s.tmem = take(16);
  s.total = o;
  return s;
}

template <int DP>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_cached_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ wqimg,
                      const uint8_t* __restrict__ woimg, const uint8_t* __restrict__ kvimg,
                      const float* __restrict__ lnw, const float* __restrict__ lnb, int64_t M, int d, int H, int hd,
                      int L, int NKV, int64_t n_tiles, float eps, const float* __restrict__ kmax,
                      long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int QN = H * HP;
#define FFD_STAMP(slot) do { if (dbg && blockIdx.x == 0 && ntile == 2 && lane == 0) dbg[slot] = clock64(); } while (0)
  const int NQ = NKV >> 2;  // key columns per softmax group (multiple of 16)
  const AttnSmem sm = attn_smem(DP, QN, NKV, d);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* k_full = bars;                       // [KV_STAGES] tx
  uint64_t* k_empty = k_full + KV_STAGES;        // [KV_STAGES] commit
  uint64_t* v_full = k_empty + KV_STAGES;        // [KV_STAGES] tx
  uint64_t* v_empty = v_full + KV_STAGES;        // [KV_STAGES] commit
  uint64_t* w_full = v_empty + KV_STAGES;        // [1] tx (Wq + Wo, once)
  uint64_t* t_full = w_full + 1;                 // [2] tx: fp32 tile landed
  uint64_t* y_ready = t_full + 2;                // [2] 512 arrivals: LayerNorm output written to the tile buffer
  uint64_t* x_full = y_ready + 2;                // [1] 512 arrivals
  uint64_t* q_full = x_full + 1;                 // [1] commit
  uint64_t* qbf_full = q_full + 1;               // [1] 512 arrivals
  uint64_t* s_full = qbf_full + 1;               // [2] commit
  uint64_t* p_full = s_full + 2;                 // [2] 512 arrivals (all four groups)
  uint64_t* o_full = p_full + 2;                 // [2] commit
  uint64_t* o_empty = o_full + 2;                // [2] 128 arrivals (the collecting group)
  uint64_t* obf_full = o_empty + 2;              // [1] 128 arrivals (the read-out warps)
  uint64_t* out_full = obf_full + 1;             // [1] commit
  uint64_t* out_empty = out_full + 1;            // [1] 512 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.tmem);
  float* consts = reinterpret_cast<float*>(smem + sm.consts);
  float* mq = reinterpret_cast<float*>(smem + sm.pmax);   // [4][4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&t_full[b], 1);
      mbar_init(&y_ready[b], 512);
    }
    mbar_init(x_full, 512);
    mbar_init(q_full, 1);
    mbar_init(qbf_full, 512);
    for (int p = 0; p < 2; ++p) {
      mbar_init(&s_full[p], 1);
      mbar_init(&p_full[p], 512);
      mbar_init(&o_full[p], 1);
      mbar_init(&o_empty[p], 128);
    }
    mbar_init(obf_full, 128);
    mbar_init(out_full, 1);
    mbar_init(out_empty, 512);
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

  if (warp == 0) {
    // ================= producer =================
    if (lane == 0) {
      mbar_expect_tx(w_full, sm.wq_bytes + sm.wo_bytes);
      bulk_load(smem + sm.wq, wqimg, sm.wq_bytes, w_full);
      bulk_load(smem + sm.wo, woimg, sm.wo_bytes, w_full);
      uint32_t nk = 0, nv = 0, ntile = 0;
      const size_t head_stride = (size_t)sm.k_bytes + sm.v_bytes;
      auto load_k = [&](int h) {
        const uint32_t st = nk % KV_STAGES, u = nk / KV_STAGES;
        mbar_wait(&k_empty[st], (u & 1) ^ 1);
        mbar_expect_tx(&k_full[st], sm.k_bytes);
        bulk_load(smem + sm.kring + st * sm.k_bytes, kvimg + h * head_stride, sm.k_bytes, &k_full[st]);
        ++nk;
      };
      auto load_v = [&](int h) {
        const uint32_t st = nv % V_STAGES, u = nv / V_STAGES;
        mbar_wait(&v_empty[st], (u & 1) ^ 1);
        mbar_expect_tx(&v_full[st], sm.v_bytes);
        bulk_load(smem + sm.vring + st * sm.v_bytes, kvimg + h * head_stride + sm.k_bytes, sm.v_bytes, &v_full[st]);
        ++nv;
      };
      auto load_tile = [&](int64_t tile, uint32_t nt) {
        const uint32_t b = nt & 1;
        const int64_t row0 = tile * TILE_M;
        const int64_t rows = M - row0 < TILE_M ? M - row0 : TILE_M;
        mbar_expect_tx(&t_full[b], (uint32_t)(rows * d * 4));
        bulk_load(smem + sm.x32 + b * sm.tile_stride, x + row0 * d, (uint32_t)(rows * d * 4), &t_full[b]);
      };
      // the producer also stores finished tiles, so no compute warp ever waits on a bulk copy
      auto store_tile = [&](int64_t tile, uint32_t nt) {
        const uint32_t b = nt & 1;
        const int64_t row0 = tile * TILE_M;
        const int64_t rows = M - row0 < TILE_M ? M - row0 : TILE_M;
        mbar_wait(&y_ready[b], (nt >> 1) & 1);
        bulk_store(y + row0 * d, smem + sm.x32 + b * sm.tile_stride, (uint32_t)(rows * d * 4));
        bulk_commit();
        bulk_wait_read();
      };
      if ((int64_t)blockIdx.x < n_tiles) load_tile(blockIdx.x, 0);
      if ((int64_t)blockIdx.x + gridDim.x < n_tiles) load_tile((int64_t)blockIdx.x + gridDim.x, 1);
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
        // consumption order of the MMA warp: K0, K1, then per head h: V_h, K_{h+2}
        load_k(0);
        if (H > 1) load_k(1);
        if (ntile > 0) {
          // previous tile: store it, then reuse its buffer for the tile after this one
          store_tile(tile - gridDim.x, ntile - 1);
          if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x, ntile + 1);
        }
        for (int h = 0; h < H; ++h) {
          load_v(h);
          if (h + 2 < H) load_k(h + 2);
        }
      }
      if (ntile > 0) {
        store_tile((int64_t)blockIdx.x + (int64_t)(ntile - 1) * gridDim.x, ntile - 1);
        bulk_wait_all();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) =================
    {
      const uint32_t idesc_q = make_idesc(TILE_M, QN);
      const uint32_t idesc_s = make_idesc(TILE_M, NKV);
      constexpr uint32_t IDESC_PV = make_idesc_f16(TILE_M, HP);
      constexpr uint32_t IDESC_OUT = make_idesc(TILE_M, DP);
      constexpr uint32_t LBO128 = 128 * 16;
      const uint32_t xo = smem_u32(smem + sm.xo), qbf = smem_u32(smem + sm.qbf);
      const uint32_t wq = smem_u32(smem + sm.wq), wo = smem_u32(smem + sm.wo);
      uint32_t nk = 0, nv = 0, ntile = 0;
      uint32_t nhead0 = 0, nhead1 = 0;  // heads processed so far per score buffer
      auto issue_s = [&](int h) {
        const int p = h & 1;
        const uint32_t st = nk % KV_STAGES;
        mbar_wait(&k_full[st], (nk / KV_STAGES) & 1);
        tc_fence_after();
        umma(tmem_base + (p ? COL_S1 : COL_S0), make_desc(qbf + h * 2 * LBO128, LBO128, 128),
             make_desc(smem_u32(smem + sm.kring + st * sm.k_bytes), NKV * 16, 128), idesc_s, 0);
        umma_commit(&s_full[p]);
        umma_commit(&k_empty[st]);
        ++nk;
      };
      auto issue_q = [&](uint32_t nt) {
        mbar_wait(x_full, nt & 1);
        tc_fence_after();
        umma_ss_chain<DP / 16>(tmem_base + COL_S1, make_desc(xo, LBO128, 128), (2 * LBO128) >> 4,
                               make_desc(wq, QN * 16, 128), (2 * QN * 16) >> 4, idesc_q, 0);
        umma_commit(q_full);
      };
      mbar_wait(w_full, 0);
      if ((int64_t)blockIdx.x < n_tiles) issue_q(0);
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
        mbar_wait(qbf_full, ntile & 1);  // Q is in shared memory as bf16; its TMEM columns are free
        mbar_wait(out_empty, (ntile & 1) ^ 1);  // S0 doubles as the previous tile's out-proj accumulator
        tc_fence_after();
        issue_s(0);
        if (H > 1) issue_s(1);
        for (int h = 0; h < H; ++h) {
          const int p = h & 1;
          const uint32_t n = p ? nhead1++ : nhead0++;
          mbar_wait(&p_full[p], n & 1);          // P_h (fp16 pairs) sits in the score buffer
          FFD_STAMP(h * 8 + 0);
          mbar_wait(&o_empty[p], (n & 1) ^ 1);   // O of this pair's previous head has been read out
          FFD_STAMP(h * 8 + 1);
          const uint32_t st = nv % V_STAGES;
          mbar_wait(&v_full[st], (nv / V_STAGES) & 1);
          tc_fence_after();
          FFD_STAMP(h * 8 + 2);
          const uint32_t vaddr = smem_u32(smem + sm.vring + st * sm.v_bytes);
          const uint32_t pcol = tmem_base + (p ? COL_S1 : COL_S0);
          // every softmax group keeps its packed P at the start of its own quarter of the buffer:
          // one chain of kq (1..3) K-steps per quarter
          const int kq = NQ >> 4;  // k-steps (16 keys) per quarter
          const uint64_t vdesc = make_desc(vaddr, HP * 16, 128);
          if (kq == 3) {
            umma_ts_4x3(tmem_base + COL_O + 64 * p, pcol, NQ, vdesc, (2 * HP * 16) >> 4, IDESC_PV);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t ocol = tmem_base + COL_O + 64 * p + 16 * q;  // quarter q has its own accumulator
              const uint32_t a0 = pcol + q * NQ;
              const uint64_t b0 = vdesc + (uint64_t)(q * kq) * ((2 * HP * 16) >> 4);
              if (kq == 2) umma_ts_chain<2>(ocol, a0, 8, b0, (2 * HP * 16) >> 4, IDESC_PV, 0);
              else umma_ts_chain<1>(ocol, a0, 8, b0, (2 * HP * 16) >> 4, IDESC_PV, 0);
            }
          }
          umma_commit(&o_full[p]);
          umma_commit(&v_empty[st]);
          ++nv;
          FFD_STAMP(h * 8 + 3);
          if (h + 2 < H) issue_s(h + 2);  // after PV_h in issue order: it overwrites P_h's columns
          FFD_STAMP(h * 8 + 4);
        }
        // ---- the next tile's Q projection goes first: it only needs the score columns of the odd heads, which
        //      the last P.V (in issue order) has finished with, and its latency hides behind this tile's epilogue ----
        if (tile + gridDim.x < n_tiles) issue_q(ntile + 1);
        // ---- output projection (A operand: the O columns the read-out warps wrote over the dead Q tiles) ----
        mbar_wait(obf_full, ntile & 1);
        tc_fence_after();
        umma_ss_chain<DP / 16>(tmem_base + COL_OUT, make_desc(qbf, LBO128, 128), (2 * LBO128) >> 4,
                               make_desc(wo, DP * 16, 128), (2 * DP * 16) >> 4, IDESC_OUT, 0);
        umma_commit(out_full);
      }
    }
  } else if (warp >= 18) {
    // ================= O read-out warps =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    // O goes where the Q tiles of already-consumed heads were: head h lands in 8-column groups <= (6h+5)/8, whose
    // Q data (heads <= h/2) the score MMAs have long finished with (SS-mode operands are read in issue order)
    uint8_t* xo = smem + sm.qbf;
    uint32_t ntile = 0, cnt0 = 0, cnt1 = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
      for (int h = 0; h < H; ++h) {
        const int b = h & 1;
        const uint32_t n = b ? cnt1++ : cnt0++;
        mbar_wait(&o_full[b], n & 1);
        tc_fence_after();
        // combine the four quarters: weights exp2(m_q - max m_q); slot 15 of each O_q is that quarter's row sum
        const float* mh = mq + ((h & 3) * 4) * TILE_M + row;
        const float q0 = mh[0], q1 = mh[TILE_M], q2 = mh[2 * TILE_M], q3 = mh[3 * TILE_M];
        const float mm = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
        const float w0 = exp2_fma(q0 - mm), w1 = exp2_fma(q1 - mm), w2 = exp2_fma(q2 - mm), w3 = exp2_fma(q3 - mm);
        float acc[HP];
        {
          uint32_t r[32];
          FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 64 * b, r);
          FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 64 * b + 16, (r + 16));
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < HP; ++c) acc[c] = w0 * __uint_as_float(r[c]) + w1 * __uint_as_float(r[16 + c]);
          FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 64 * b + 32, r);
          FFD_TMEM_LD16(tmem_base + lane_base + COL_O + 64 * b + 48, (r + 16));
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < HP; ++c)
            acc[c] = fmaf(w2, __uint_as_float(r[c]), fmaf(w3, __uint_as_float(r[16 + c]), acc[c]));
        }
        tc_fence_before();
        mbar_arrive(&o_empty[b]);
        const float inv = rcp_fma(acc[HP - 1]);
#pragma unroll
        for (int c = 0; c < HP - 1; ++c) {
          if (c < hd) {
            const int col = h * hd + c;
            *reinterpret_cast<__nv_bfloat16*>(xo + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                __float2bfloat16_rn(acc[c] * inv);
          }
        }
      }
      // constant columns of the out-proj operand: d, d+1 = 1 (bias hi/lo), the rest of the padding = 0
      for (int c = d; c < DP; ++c)
        *reinterpret_cast<__nv_bfloat16*>(xo + row_off + (c >> 3) * (TILE_M * 16) + (c & 7) * 2) =
            __float2bfloat16_rn((c == d || c == d + 1) ? 1.f : 0.f);
      fence_proxy_async();
      mbar_arrive(obf_full);
    }
  } else {
    // ================= softmax / epilogue groups =================
    const int wg = (warp - 2) >> 2;        // 0..3: quarter of the key columns
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    uint8_t* xo = smem + sm.xo;
    uint8_t* qbf = smem + sm.qbf;
    uint32_t ntile = 0;
    uint32_t cnt0 = 0, cnt1 = 0;  // uses so far of score buffer / O buffer 0 and 1
    // stage this group's share of X (k-groups of 8 columns) of tile `tile` (the nt-th of this CTA) as bf16
    auto stage_x = [&](int64_t tile, uint32_t nt) {
      const uint32_t tb = nt & 1;
      const float4* src = reinterpret_cast<const float4*>(smem + sm.x32 + tb * sm.tile_stride) + (size_t)row * (d / 4);
      const bool valid = tile * TILE_M + row < M;
      mbar_wait(&t_full[tb], (nt >> 1) & 1);
      constexpr int NG = DP / 8;
      for (int kg = wg; kg < NG; kg += 4) {
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
        q.x = pack_bf16(a.x, a.y);
        q.y = pack_bf16(a.z, a.w);
        q.z = pack_bf16(b4.x, b4.y);
        q.w = pack_bf16(b4.z, b4.w);
        *reinterpret_cast<uint4*>(xo + row_off + kg * (TILE_M * 16)) = q;
      }
      fence_proxy_async();
      mbar_arrive(x_full);
    };
    if ((int64_t)blockIdx.x < n_tiles) stage_x(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ntile) {
      const int64_t row0 = tile * TILE_M;
      const bool valid = row0 + row < M;
      const uint32_t tb = ntile & 1;
      float* tile32 = reinterpret_cast<float*>(smem + sm.x32 + tb * sm.tile_stride);
      float4* myrow = reinterpret_cast<float4*>(tile32 + row * d);
      // ---- Q: TMEM fp32 -> bf16 A-operand tiles, heads split over the four groups ----
      if (warp == 2) FFD_STAMP(303);
      mbar_wait(q_full, ntile & 1);
      tc_fence_after();
      if (warp == 2) FFD_STAMP(304);
      for (int h = wg; h < H; h += 4) {
        uint32_t r[16];
        FFD_TMEM_LD16(tmem_base + lane_base + COL_S1 + h * HP, r);
        tmem_ld_wait();
        uint4 q0, q1;
        q0.x = pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1]));
        q0.y = pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3]));
        q0.z = pack_bf16(__uint_as_float(r[4]), __uint_as_float(r[5]));
        q0.w = pack_bf16(__uint_as_float(r[6]), __uint_as_float(r[7]));
        q1.x = pack_bf16(__uint_as_float(r[8]), __uint_as_float(r[9]));
        q1.y = pack_bf16(__uint_as_float(r[10]), __uint_as_float(r[11]));
        q1.z = pack_bf16(__uint_as_float(r[12]), __uint_as_float(r[13]));
        q1.w = pack_bf16(__uint_as_float(r[14]), __uint_as_float(r[15]));
        *reinterpret_cast<uint4*>(qbf + row_off + (2 * h) * (TILE_M * 16)) = q0;
        *reinterpret_cast<uint4*>(qbf + row_off + (2 * h + 1) * (TILE_M * 16)) = q1;
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(qbf_full);
      if (warp == 2) FFD_STAMP(305);
      // ---- heads ----
      for (int h = 0; h < H; ++h) {
        const int b = h & 1;
        const uint32_t n = b ? cnt1++ : cnt0++;
        const uint32_t scol = tmem_base + lane_base + (b ? COL_S1 : COL_S0) + wg * NQ;  // this group's columns
        mbar_wait(&s_full[b], n & 1);
        tc_fence_after();
        if (warp == 2) FFD_STAMP(128 + h * 8 + 0);
        if (warp == 2) FFD_STAMP(128 + h * 8 + 1);
        // p = exp2(s - m_q) as fp16 pairs with m_q the row maximum over THIS group's quarter of the keys, written
        // in place: 16 score columns [c, c+16) -> 8 packed columns [c/2, c/2+8) of the quarter.  Padded keys need no
        // mask: slot 15 of every Q head is the constant 1 and slot 15 of a padded K row is -29952 (a quarter made
        // only of padding gets m_q = -29952 and weight exp2(m_q - max) = 0 in the combine).
        uint32_t r[48];
        FFD_TMEM_LD16(scol, r);
        if (NQ > 16) FFD_TMEM_LD16(scol + 16, (r + 16));
        if (NQ > 32) FFD_TMEM_LD16(scol + 32, (r + 32));
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 16; e += 2) m0 = max3(m0, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
        if (NQ > 16) {
#pragma unroll
          for (int e = 16; e < 32; e += 2) m1 = max3(m1, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
        }
        if (NQ > 32) {
#pragma unroll
          for (int e = 32; e < 48; e += 2) m2 = max3(m2, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
        }
        const float mb = max3(m0, m1, m2);
        mq[((h & 3) * 4 + wg) * TILE_M + row] = mb;
        if (warp == 2) FFD_STAMP(128 + h * 8 + 2);
        {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2)
            pk[e >> 1] = ((POLY_PAIRS >> (e >> 1)) & 1)
                             ? ex2_pack_f16_poly(__uint_as_float(r[e]) - mb, __uint_as_float(r[e + 1]) - mb)
                             : ex2_pack_f16(__uint_as_float(r[e]) - mb, __uint_as_float(r[e + 1]) - mb);
          FFD_TMEM_ST8(scol, pk);
          if (NQ > 16) {
#pragma unroll
            for (int e = 0; e < 16; e += 2)
              pk[e >> 1] = ((POLY_PAIRS >> (e >> 1)) & 1)
                               ? ex2_pack_f16_poly(__uint_as_float(r[16 + e]) - mb, __uint_as_float(r[17 + e]) - mb)
                               : ex2_pack_f16(__uint_as_float(r[16 + e]) - mb, __uint_as_float(r[17 + e]) - mb);
            FFD_TMEM_ST8(scol + 8, pk);
          }
          if (NQ > 32) {
#pragma unroll
            for (int e = 0; e < 16; e += 2)
              pk[e >> 1] = ((POLY_PAIRS >> (e >> 1)) & 1)
                               ? ex2_pack_f16_poly(__uint_as_float(r[32 + e]) - mb, __uint_as_float(r[33 + e]) - mb)
                               : ex2_pack_f16(__uint_as_float(r[32 + e]) - mb, __uint_as_float(r[33 + e]) - mb);
            FFD_TMEM_ST8(scol + 16, pk);
          }
        }
        if (warp == 2) FFD_STAMP(128 + h * 8 + 3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[b]);
        if (warp == 2) FFD_STAMP(128 + h * 8 + 4);
      }
      // the next tile's X can be staged now (this tile's Q projection is long done with the buffer), so that its
      // Q projection runs while this tile's out-proj / LayerNorm finish
      if (tile + gridDim.x < n_tiles) stage_x(tile + gridDim.x, ntile + 1);
      if (warp == 2) FFD_STAMP(306);
      // ---- out-proj result + residual -> LayerNorm -> tile buffer -> bulk store.  All four groups share the
      //      work: 16-column unit u of the accumulator belongs to group u & 3; the row sums are exchanged
      //      through shared memory. ----
      {
        mbar_wait(out_full, ntile & 1);
        tc_fence_after();
        if (warp == 10) FFD_STAMP(307);
        constexpr int NU = DP / 16;            // 16-column units
        constexpr int MAXU = (NU + 3) / 4;     // units per group
        float vals[MAXU][16];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXU; ++i) {
          const int u = wg + 4 * i;
          if (u < NU) {
            uint32_t r[16];
            FFD_TMEM_LD16(tmem_base + lane_base + COL_OUT + 16 * u, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = 16 * u + 4 * j;
              float4 xr = make_float4(0.f, 0.f, 0.f, 0.f);
              if (c < d && valid) xr = myrow[c >> 2];
              const float a0 = (c < d) ? __uint_as_float(r[4 * j]) + xr.x : 0.f;
              const float a1 = (c < d) ? __uint_as_float(r[4 * j + 1]) + xr.y : 0.f;
              const float a2 = (c < d) ? __uint_as_float(r[4 * j + 2]) + xr.z : 0.f;
              const float a3 = (c < d) ? __uint_as_float(r[4 * j + 3]) + xr.w : 0.f;
              vals[i][4 * j] = a0, vals[i][4 * j + 1] = a1, vals[i][4 * j + 2] = a2, vals[i][4 * j + 3] = a3;
              s1 += (a0 + a1) + (a2 + a3);
              s2 = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2))));
            }
          }
        }
        tc_fence_before();
        mbar_arrive(out_empty);
        float2* ps = reinterpret_cast<float2*>(mq);  // the quarter-maxima buffer is idle between tiles
        ps[wg * TILE_M + row] = make_float2(s1, s2);
        named_bar_sync(5, 512);
        const float2 p0 = ps[row], p1 = ps[TILE_M + row], p2 = ps[2 * TILE_M + row], p3 = ps[3 * TILE_M + row];
        const float mean = ((p0.x + p1.x) + (p2.x + p3.x)) / (float)d;
        const float ex2v = ((p0.y + p1.y) + (p2.y + p3.y)) / (float)d;
        const float rstd = rsqrtf(fmaxf(ex2v - mean * mean, 0.f) + eps);
        if (valid) {
#pragma unroll
          for (int i = 0; i < MAXU; ++i) {
            const int u = wg + 4 * i;
            if (u < NU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = 16 * u + 4 * j;
                if (c < d) {
                  float4 o;
                  o.x = (vals[i][4 * j] - mean) * rstd * consts[c] + consts[DP + c];
                  o.y = (vals[i][4 * j + 1] - mean) * rstd * consts[c + 1] + consts[DP + c + 1];
                  o.z = (vals[i][4 * j + 2] - mean) * rstd * consts[c + 2] + consts[DP + c + 2];
                  o.w = (vals[i][4 * j + 3] - mean) * rstd * consts[c + 3] + consts[DP + c + 3];
                  myrow[c >> 2] = o;
                }
              }
            }
          }
        }
        if (warp == 10) FFD_STAMP(308);
        fence_proxy_async();
        mbar_arrive(&y_ready[tb]);
      }
    }

  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- packing ---------------------------------------------------------------------------------------
// WqP [QN n x DP k]: row h*16+c = Wq[h*hd+c, :] * log2(e)/sqrt(hd), bias hi/lo in columns d, d+1.
// WoP [DP n x DP k]: row n = Wo[n, :], bias hi/lo in columns d, d+1.   bf16 core-matrix order.
__global__ void pack_attn_kernel(const float* __restrict__ in_w, const float* __restrict__ in_b,
                                 const float* __restrict__ out_w, const float* __restrict__ out_b,
                                 __nv_bfloat16* __restrict__ wq, __nv_bfloat16* __restrict__ wo, int d, int H, int DP) {
  const int hd = d / H, QN = H * HP;
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
      const int h = n / HP, c = n % HP;
      if (c == HP - 1) {
        v = (k == d) ? 1.f : 0.f;  // constant-1 slot of every Q head (pairs with the key-padding slot of K)
      } else if (c < hd) {
        const int r = h * hd + c;
        const float b = in_b[r] * scale;
        if (k < d) v = in_w[r * d + k] * scale;
        else if (k == d) v = __bfloat162float(__float2bfloat16_rn(b));
        else if (k == d + 1) v = b - __bfloat162float(__float2bfloat16_rn(b));
      }
    } else if (n < d) {
      const float b = out_b[n];
      if (k < d) v = out_w[n * d + k];
      else if (k == d) v = __bfloat162float(__float2bfloat16_rn(b));
      else if (k == d + 1) v = b - __bfloat162float(__float2bfloat16_rn(b));
    }
    (is_o ? wo : wq)[e] = __float2bfloat16_rn(v);
  }
}

// K image per (layer, head): B operand [NKV n(key) x 16 k(dim)], bf16; slot 15 of a padded key row is
// -29952 (slot 15 of Q is the constant 1).  V image: B operand [16 n(dim) x NKV k(key)], fp16 (P is fp16),
// with row 15 set to 1 for the valid keys so that column 15 of O = P . V^T is the softmax row sum.
__global__ void pack_kv_kernel(const float* __restrict__ kc, const float* __restrict__ vc,
                               uint16_t* __restrict__ img, int n_lh, int L, int hd, int NKV) {
  const int per = NKV * HP;  // elements per image
  const int total = n_lh * 2 * per;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lh = i / (2 * per);
    const int r = i % (2 * per);
    const bool is_v = r >= per;
    const int e = is_v ? r - per : r;
    const int rows = is_v ? HP : NKV;
    const int kg = e / (rows * 8), rem = e % (rows * 8);
    const int n = (rem / 64) * 8 + (rem % 64) / 8, k = kg * 8 + rem % 8;
    const int key = is_v ? k : n, dim = is_v ? n : k;
    float v = 0.f;
    if (key < L && dim < hd) v = (is_v ? vc : kc)[((int64_t)lh * L + key) * hd + dim];
    if (is_v && key < L && dim == HP - 1) v = 1.f;        // O[:, 15] = softmax row sum
    if (!is_v && key >= L && dim == HP - 1) v = -29952.f;  // padded keys: score = -29952 -> weight 0
    img[i] = is_v ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// Kmax[l, h] = max_j |bf16(k_j)|_2 over the valid keys: the softmax shift bound is |q| Kmax (one warp per head)
__global__ void kmax_kernel(const float* __restrict__ kc, float* __restrict__ out, int L, int hd) {
  const int lh = blockIdx.x, lane = threadIdx.x;
  float mx = 0.f;
  for (int j = lane; j < L; j += 32) {
    float n2 = 0.f;
    for (int c = 0; c < hd; ++c) {
      const float v = __bfloat162float(__float2bfloat16_rn(kc[((int64_t)lh * L + j) * hd + c]));
      n2 = fmaf(v, v, n2);
    }
    mx = fmaxf(mx, n2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) out[lh] = sqrtf(mx);
}

}  // namespace tc

bool tc_attn_supported(const ffd_dims& m) {
  const int hd = m.d_model / m.n_head;
  return tc_supported(m) && hd < tc::HP && m.n_head * tc::HP <= 192 && m.max_len <= 192;
}

static int nkv_of(const ffd_dims& m) { return (m.max_len + 63) / 64 * 64; }

int64_t tc_attn_image_bytes(const ffd_dims& m) {
  const int DP = tc::dp_of(m.d_model), QN = m.n_head * tc::HP;
  return (int64_t)(QN * DP + DP * DP) * 2;
}

static int64_t kv_images_only_bytes(const ffd_dims& m) {
  return (int64_t)m.n_layers * m.n_head * 2 * nkv_of(m) * tc::HP * 2;
}

// K/V operand images of all layers and heads, then Kmax[NL, H] (fp32)
int64_t tc_kv_image_bytes(const ffd_dims& m) { return kv_images_only_bytes(m) + (int64_t)m.n_layers * m.n_head * 4; }

int tc_pack_attn(const ffd_dims& m, const float* weights, void* img_base, int64_t layer_stride_bytes,
                 int64_t attn_off_bytes, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const int DP = tc::dp_of(m.d_model), QN = m.n_head * tc::HP;
  for (int l = 0; l < m.n_layers; ++l) {
    const float* w = weights + ar.layer0 + (int64_t)l * ar.layer_stride;
    __nv_bfloat16* wq = (__nv_bfloat16*)((uint8_t*)img_base + l * layer_stride_bytes + attn_off_bytes);
    tc::pack_attn_kernel<<<32, 256, 0, s>>>(w + ar.in_w, w + ar.in_b, w + ar.out_w, w + ar.out_b, wq,
                                            wq + (size_t)QN * DP, m.d_model, m.n_head, DP);
    FFD_CHECK_LAUNCH("pack_attn_kernel");
  }
  return FFD_OK;
}

int tc_pack_kv(const ffd_dims& m, const float* kc, const float* vc, void* kvimg, cudaStream_t s) {
  const int n_lh = m.n_layers * m.n_head;
  tc::pack_kv_kernel<<<148, 256, 0, s>>>(kc, vc, (uint16_t*)kvimg, n_lh, m.max_len, m.d_model / m.n_head,
                                         nkv_of(m));
  FFD_CHECK_LAUNCH("pack_kv_kernel");
  tc::kmax_kernel<<<n_lh, 32, 0, s>>>(kc, (float*)((uint8_t*)kvimg + kv_images_only_bytes(m)), m.max_len,
                                      m.d_model / m.n_head);
  FFD_CHECK_LAUNCH("kmax_kernel");
  return FFD_OK;
}

long long* g_attn_dbg = nullptr;  // optional device buffer of clock64 stamps (profiling aid)

template <int DP>
static int launch_attn(const ffd_dims& m, const uint8_t* wqimg, const uint8_t* kvimg_layer, const float* kmax,
                       const float* lnw, const float* lnb, const float* x, float* y, int64_t M, cudaStream_t s) {
  static int n_sm = 0;
  const int QN = m.n_head * tc::HP, NKV = nkv_of(m);
  const tc::AttnSmem sm = tc::attn_smem(DP, QN, NKV, m.d_model);
  if (n_sm == 0) {
    int dev = 0;
    FFD_CHECK_CUDA(cudaGetDevice(&dev));
    FFD_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::attn_cached_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      sm.total));
  const int64_t n_tiles = (M + tc::TILE_M - 1) / tc::TILE_M;
  const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
  tc::attn_cached_tc_kernel<DP><<<grid, tc::ATT_THREADS, sm.total, s>>>(
      x, y, wqimg, wqimg + (size_t)QN * DP * 2, kvimg_layer, lnw, lnb, M, m.d_model, m.n_head,
      m.d_model / m.n_head, m.max_len, NKV, n_tiles, 1e-5f, kmax, g_attn_dbg);
  FFD_CHECK_LAUNCH("attn_cached_tc_kernel");
  return FFD_OK;
}

int tc_attn_cached(const ffd_dims& m, const void* attn_img, const void* kvimg, const float* weights, int layer,
                   const float* x, float* y, int64_t M, cudaStream_t s) {
  const Arena ar = arena_of(m);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const uint8_t* kvl = (const uint8_t*)kvimg + (size_t)layer * m.n_head * 2 * nkv_of(m) * tc::HP * 2;
  const uint8_t* wq = (const uint8_t*)attn_img;
  const float* km = (const float*)((const uint8_t*)kvimg + kv_images_only_bytes(m)) + (size_t)layer * m.n_head;
  switch (tc::dp_of(m.d_model)) {
    case 16: return launch_attn<16>(m, wq, kvl, km, w + ar.n1w, w + ar.n1b, x, y, M, s);
    case 32: return launch_attn<32>(m, wq, kvl, km, w + ar.n1w, w + ar.n1b, x, y, M, s);
    case 48: return launch_attn<48>(m, wq, kvl, km, w + ar.n1w, w + ar.n1b, x, y, M, s);
    case 64: return launch_attn<64>(m, wq, kvl, km, w + ar.n1w, w + ar.n1b, x, y, M, s);
    case 80: return launch_attn<80>(m, wq, kvl, km, w + ar.n1w, w + ar.n1b, x, y, M, s);
  }
  set_error("tensor-core attention: unsupported d_model %d", m.d_model);
  return FFD_ERR_UNSUPPORTED;
}

}  // namespace ffd
