// ONE persistent tcgen05 / TMEM kernel per PURE-CACHE encoder layer (998 of 1000 default steps):
//   y1 = LN1(x + out_proj(softmax(q K_c^T / sqrt(hd)) V_c)),  q = x Wq^T + bq      cached_transformer.py:228-258,309-321
//   y  = LN2(y1 + W2 relu(W1 y1 + b1) + b2)                                        cached_transformer.py:130-132,325-327
// The attention block is bound by MUFU.EX2 (one exponential per token x head x key: 16 lanes/clk/SM; tensor pipe
// 12 % busy) and the FFN block by the tensor pipe (MUFU idle), so as two kernels (ffd_tc_attn.cu + ffd_tc.cu) a layer
// costs the SUM of the two bounds.  Here one CTA per SM runs both at once, software-pipelined over 128-token tiles:
// while the softmax warps work on the heads of tile p, the tensor pipe runs the FFN of tile p-1 -- and the [M, d]
// fp32 round trip between the two blocks is gone (y1 goes from registers straight into TMEM).
//
// Per tile (period p = the heads of tile p are running):
//   attention heads of tile p     "virtual head" n = (head, chunk of 96 keys).  Two softmax TEAMS of 8 warps ping-pong
//                                 over the virtual heads (team = n & 1), each with its own 96-column score buffer:
//                                 S_n = Qtile . K_n^T (one MMA, N = 96) -> the team's warps load S_n, take the row
//                                 maximum of their 48 columns, write P_n = exp2(S_n - m) as fp16 pairs IN PLACE over
//                                 their own score columns -> O_n = P_n . [V_n | 1] (two 3-step chains, one accumulator
//                                 per 48-key group; the ones row of V returns the row sum) -> the same warp issues
//                                 S_{n+2} right behind it (in-order tensor pipe: no round trip).  While one team
//                                 waits for that turn-around the other one keeps the MUFU pipe busy.  Read-out warps
//                                 combine the partial (O, rowsum, max) of a head (flash-style, exact) into the bf16
//                                 operand of the out-projection.
//   tile warps, in this order     F1(p-1): out-projection (scratch MMA) + residual + LayerNorm1 -> y1: bf16 into the
//                                 TMEM A operand of the FFN, and y1 + b2 (fp32) into the FFN's Y ACCUMULATOR, so that
//                                 GEMM2 adds the FFN on top of the residual and no fp32 copy of y1 is kept anywhere;
//                                 Q(p+1): Q projection of the next tile (scratch MMA) -> bf16 Q tiles;
//                                 FFN(p-1) starts;  X(p+2): next-but-one tile read from global memory (coalesced) into
//                                 the bf16 operand of its Q projection;  F2(p-1): Y -> LayerNorm2 -> global memory.
//   FFN of tile p-1               32 half-slices of 64 hidden units, H double-buffered in TMEM: GEMM1 H_b = X . W1_j^T
//                                 (A from TMEM) -> epilogue warps relu + bf16 in place -> GEMM2 Y += relu(H_b) . W2_j^T;
//                                 weight half-slices (10 KB) stream through two 2-stage rings.
// The scratch MMAs (out-projection, Q projection) use the H columns while the FFN is between tiles.
//
// TMEM (512 columns): S/P team 0 [0, 96) | S/P team 1 [96, 192) | O 4 x 16 [192, 256) | X [256, 304) | Y [304, 384) |
//                     H0 [384, 448) | H1 [448, 512)
// Warp roles (32 warps; registers per warp group via setmaxnreg: 72 / 64 / 72 / 48 / 40):
//   0..15 softmax: team w>>3, key group (w>>2)&1, lane quadrant w&3      16..19 FFN epilogue (relu)
//   20..23 tile warps (warp 20 issues the scratch MMAs and the residual-tile loads)      24..27 O read-out
//   28 K/V image producer   29 attention MMA issuer (P.V of head n, then S of head n+2)
//   30 FFN weight producer  31 FFN MMA issuer          (highest ids = highest issue priority: see the kernel)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ffd_tc_ptx.cuh"

namespace ffd {
namespace tc {

constexpr int LAYER_THREADS = 1024;
constexpr int VK = 96;           // keys per virtual head
constexpr int LHP = 16;          // slots per Q tile (UMMA K of the score MMA, N of the P.V MMA)
constexpr int L_KSTAGES = 4;     // K image ring
constexpr int L_VSTAGES = 3;     // V image ring
constexpr int FCH = 64;          // hidden units per FFN half-slice
constexpr int L_WSTAGES = 2;     // W1 / W2 half-slice rings

constexpr uint32_t LC_S = 0;      // + 96 * team
constexpr uint32_t LC_O = 192;    // + 32 * team + 16 * group
constexpr uint32_t LC_X = 256;
constexpr uint32_t LC_Y = 304;
constexpr uint32_t LC_H = 384;    // + 64 * buffer; scratch passes use [384, 512) while the FFN is between tiles

struct LayerSmem {
  int res, xo, obf, qbf, wqo, kring, vring, w1, w2, stage, consts, mq, bars, tmem, total;
  int tile_bytes, k_bytes, v_bytes, wq_bytes, wo_bytes, qbf_bytes, w1_bytes, w2_bytes;
};

__host__ __device__ inline LayerSmem layer_smem(int DP, int QN, int d) {
  LayerSmem s;
  int o = 0;
  auto take = [&](int bytes) {
    int r = o;
    o += (bytes + 127) / 128 * 128;
    return r;
  };
  s.tile_bytes = TILE_M * d * 4;
  s.k_bytes = VK * LHP * 2;
  s.v_bytes = LHP * VK * 2;
  s.wq_bytes = QN * DP * 2;
  s.wo_bytes = DP * DP * 2;
  s.qbf_bytes = TILE_M * QN * 2;
  s.w1_bytes = FCH * DP * 2;          // bf16 [64 n x DP k]
  s.w2_bytes = DP * FCH * 2;          // bf16 [DP n x 64 k]
  s.res = take(s.tile_bytes);         // fp32 rows of the tile whose attention block is being finished (residual)
  s.xo = take(TILE_M * DP * 2);       // bf16 A operand of the Q projection
  s.stage = take(4 * 32 * 80);        // F2's store staging: per tile warp 32 rows x 16 columns (row stride 80 B)
  s.obf = take(TILE_M * DP * 2);      // bf16 A operand of the out-projection
  s.qbf = take(2 * s.qbf_bytes);      // bf16 Q tiles of this tile and the next
  s.wqo = take(s.wq_bytes > s.wo_bytes ? s.wq_bytes : s.wo_bytes);  // Wq and Wo take turns (one MMA each per tile)
  s.kring = take(L_KSTAGES * s.k_bytes);
  s.vring = take(L_VSTAGES * s.v_bytes);
  s.w1 = take(L_WSTAGES * s.w1_bytes);
  s.w2 = take(L_WSTAGES * s.w2_bytes);
  s.consts = take(5 * DP * 4);        // ln1 w, ln1 b, b2, ln2 w, ln2 b
  s.mq = take(4 * 2 * TILE_M * 4);    // group row maxima [virtual head & 3][group][row]
  s.bars = take(48 * 8);
  s.tmem = take(16);
  s.total = o;
  return s;
}

template <int DP, int D_MODEL>
__global__ void __launch_bounds__(LAYER_THREADS, 1)
layer_cached_tc_kernel(const float* __restrict__ x, float* __restrict__ y /* may alias x */,
                       const uint8_t* __restrict__ wqimg, const uint8_t* __restrict__ woimg,
                       const uint8_t* __restrict__ kvimg, const uint8_t* __restrict__ w1img,
                       const uint8_t* __restrict__ w2img, const float* __restrict__ ln1w,
                       const float* __restrict__ ln1b, const float* __restrict__ b2, const float* __restrict__ ln2w,
                       const float* __restrict__ ln2b, int64_t M, int H, int hd, int G, int QN, int NCH, int n_slices,
                       int64_t n_tiles, float eps, long long* __restrict__ dbg) {
#ifdef FFD_LAYER_STAMPS  // clock64 timeline of CTA 0 around its local tile 2 (tools/layer_stamps.py)
#define LSTAMP(cond, slot) do { if (dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (cond)) dbg[slot] = clock64(); } while (0)
#else
#define LSTAMP(cond, slot) do { } while (0)
#endif
  constexpr int d = D_MODEL;
  static_assert(D_MODEL % 8 == 0 && D_MODEL + 2 <= DP && D_MODEL + 16 > DP && DP % 16 == 0 && DP <= 80, "d_model / DP");
  extern __shared__ __align__(1024) uint8_t smem[];
  const LayerSmem sm = layer_smem(DP, QN, d);
  const int NV = H * NCH;  // virtual heads per tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* k_full = bars;                        // [3] tx
  uint64_t* k_empty = k_full + L_KSTAGES;         // [3] commit
  uint64_t* v_full = k_empty + L_KSTAGES;         // [2] tx
  uint64_t* v_empty = v_full + L_VSTAGES;         // [2] commit
  uint64_t* w_full = v_empty + L_VSTAGES;         // [1] tx: Wq or Wo landed (they alternate in one buffer)
  uint64_t* f1_full = w_full + 1;                 // [2] tx: W1 half-slice
  uint64_t* f1_empty = f1_full + L_WSTAGES;       // [2] commit
  uint64_t* f2_full = f1_empty + L_WSTAGES;       // [2] tx: W2 half-slice
  uint64_t* f2_empty = f2_full + L_WSTAGES;       // [2] commit
  uint64_t* t_full = f2_empty + L_WSTAGES;        // [1] tx: residual tile landed
  uint64_t* qbf_full = t_full + 1;                // [2] 128 arrivals: Q tiles of a token tile written
  uint64_t* s_full = qbf_full + 2;                // [2] commit (per team)
  uint64_t* p_full = s_full + 2;                  // [2] 256 arrivals (per team)
  uint64_t* o_full = p_full + 2;                  // [2] commit (per team)
  uint64_t* o_empty = o_full + 2;                 // [2] 128 arrivals (read-out warps)
  uint64_t* obf_full = o_empty + 2;               // [1] 128 arrivals (read-out warps)
  uint64_t* obf_empty = obf_full + 1;             // [1] commit: the out-projection has read its operand
  uint64_t* scr_full = obf_empty + 1;             // [1] commit: a scratch pass has been computed
  uint64_t* x_full = scr_full + 1;                // [1] 128 arrivals: X and Y of a tile are in TMEM, scratch is free
  uint64_t* h_full = x_full + 1;                  // [2] commit
  uint64_t* r_full = h_full + 2;                  // [2] 128 arrivals: relu(H) packed in place
  uint64_t* y_full = r_full + 2;                  // [1] commit: all MMAs of a tile's FFN are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.tmem);
  float* consts = reinterpret_cast<float*>(smem + sm.consts);
  float* mq = reinterpret_cast<float*>(smem + sm.mq);  // [4][2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The warp scheduler favours the highest warp ids of a sub-partition (measured: with the MMA-issuing warps at ids
  // 1 and 3 below sixteen always-eligible softmax warps, every tcgen05.mma took ~100-150 cycles to issue and both
  // pipelines ran at the issue rate of those two warps).  So: softmax warps lowest, the latency-critical single
  // warps highest.
  constexpr int W_SOFTMAX = 0, W_EPI = 16, W_TILE = 20, W_READ = 24, W_KV = 28, W_AMMA = 29, W_FW = 30, W_FMMA = 31;
  const uint32_t my_tiles = (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // >= 1
  const uint32_t n_heads = my_tiles * (uint32_t)NV;
  auto tile_row0 = [&](uint32_t t) { return ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M; };
  auto tile_rows = [&](uint32_t t) {
    const int64_t r0 = tile_row0(t);
    return (int)(M - r0 < TILE_M ? M - r0 : TILE_M);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < L_KSTAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < L_VSTAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < L_WSTAGES; ++i) {
      mbar_init(&f1_full[i], 1);
      mbar_init(&f1_empty[i], 1);
      mbar_init(&f2_full[i], 1);
      mbar_init(&f2_empty[i], 1);
    }
    mbar_init(t_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&qbf_full[b], 128);
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 256);
      mbar_init(&o_full[b], 1);
      mbar_init(&o_empty[b], 128);
      mbar_init(&h_full[b], 1);
      mbar_init(&r_full[b], 128);
    }
    mbar_init(obf_full, 128);
    mbar_init(obf_empty, 1);
    mbar_init(scr_full, 1);
    mbar_init(x_full, 128);
    mbar_init(y_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 5 * DP; i += LAYER_THREADS) {
    const int w = i / DP, c = i % DP;
    const float* src = w == 0 ? ln1w : w == 1 ? ln1b : w == 2 ? b2 : w == 3 ? ln2w : ln2b;
    consts[i] = c < d ? src[c] : 0.f;
  }
  if (warp == W_AMMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // provably warp-uniform

  if (warp >= W_KV) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == W_KV) {
      // ================= K/V image producer (+ the first Wq) =================
      if (lane == 0) {
        mbar_expect_tx(w_full, sm.wq_bytes);
        bulk_load(smem + sm.wqo, wqimg, sm.wq_bytes, w_full);
        const size_t img_stride = (size_t)sm.k_bytes + sm.v_bytes;
        uint32_t nk = 0, nv = 0;
        auto load_k = [&](uint32_t n) {
          const uint32_t st = nk % L_KSTAGES, u = nk / L_KSTAGES;
          mbar_wait_relaxed(&k_empty[st], (u & 1) ^ 1, 200);
          mbar_expect_tx(&k_full[st], sm.k_bytes);
          bulk_load(smem + sm.kring + st * sm.k_bytes, kvimg + (size_t)(n % NV) * img_stride, sm.k_bytes, &k_full[st]);
          ++nk;
        };
        auto load_v = [&](uint32_t n) {
          const uint32_t st = nv % L_VSTAGES, u = nv / L_VSTAGES;
          mbar_wait_relaxed(&v_empty[st], (u & 1) ^ 1, 200);
          mbar_expect_tx(&v_full[st], sm.v_bytes);
          bulk_load(smem + sm.vring + st * sm.v_bytes, kvimg + (size_t)(n % NV) * img_stride + sm.k_bytes, sm.v_bytes,
                    &v_full[st]);
          ++nv;
        };
        // the K images run two virtual heads ahead of the V images (S_{n+2} is issued behind P.V_n)
        load_k(0);
        if (n_heads > 1) load_k(1);
        for (uint32_t n = 0; n < n_heads; ++n) {
          load_v(n);
          if (n + 2 < n_heads) load_k(n + 2);
        }
      }
    } else if (warp == W_FW) {
      // ================= FFN weight producer: W1_0, W1_1, then W2_j, W1_{j+2} (the order the issuer consumes) =====
      if (lane == 0) {
        uint32_t k1 = 0, k2 = 0;
        auto load_w1 = [&](int j) {
          const uint32_t st = k1 % L_WSTAGES, u = k1 / L_WSTAGES;
          mbar_wait_relaxed(&f1_empty[st], (u & 1) ^ 1, 100);
          mbar_expect_tx(&f1_full[st], sm.w1_bytes);
          bulk_load(smem + sm.w1 + st * sm.w1_bytes, w1img + (size_t)j * sm.w1_bytes, sm.w1_bytes, &f1_full[st]);
          ++k1;
        };
        auto load_w2 = [&](int j) {
          const uint32_t st = k2 % L_WSTAGES, u = k2 / L_WSTAGES;
          mbar_wait_relaxed(&f2_empty[st], (u & 1) ^ 1, 100);
          mbar_expect_tx(&f2_full[st], sm.w2_bytes);
          bulk_load(smem + sm.w2 + st * sm.w2_bytes, w2img + (size_t)j * sm.w2_bytes, sm.w2_bytes, &f2_full[st]);
          ++k2;
        };
        for (uint32_t t = 0; t < my_tiles; ++t) {
          load_w1(0);
          if (n_slices > 1) load_w1(1);
          for (int j = 0; j < n_slices; ++j) {
            load_w2(j);
            if (j + 2 < n_slices) load_w1(j + 2);
          }
        }
      }
    } else if (warp == W_AMMA) {
      // ================= attention MMA issuer (whole warp converged, one elected lane issues) =================
      // P.V of virtual head n, then -- in the same instruction stream, so that it simply runs behind it on the
      // in-order tensor pipe -- the scores of virtual head n + 2 into the buffer P_n is being read from.
      const uint32_t idesc_s = make_idesc(TILE_M, VK);
      constexpr uint32_t IDESC_PV = make_idesc_f16(TILE_M, LHP);
      constexpr uint32_t LBO128 = 128 * 16;
      const uint32_t qbf = smem_u32(smem + sm.qbf);
      const uint32_t kring = smem_u32(smem + sm.kring);
      const uint32_t vring = smem_u32(smem + sm.vring);
      constexpr uint32_t VSTEP = (2 * LHP * 16) >> 4;  // 16 keys of the V image, in 16-byte units
      const uint32_t bar0 = smem_u32(bars);
      auto bar = [&](const uint64_t* b) { return bar0 + (uint32_t)((b - bars) * 8); };
      // position of the NEXT score MMA to be issued (tile, virtual head in it, head, chunk, Q tile, place in it):
      // plain counters, so that every MMA operand is provably warp-uniform (a per-thread "waterfall" around each
      // tcgen05 instruction otherwise)
      uint32_t s_n = 0, s_t = 0, s_v = 0, s_c = 0, s_g = 0, s_qt = 0;
      auto s_advance = [&]() {
        ++s_n;
        if (++s_c == (uint32_t)NCH) {
          s_c = 0;
          if (++s_g == (uint32_t)G) s_g = 0, ++s_qt;
        }
        if (++s_v == (uint32_t)NV) s_v = 0, s_c = 0, s_g = 0, s_qt = 0, ++s_t;
      };
      auto q_desc = [&]() { return make_desc(qbf + (s_t & 1) * sm.qbf_bytes + s_qt * 2 * LBO128, LBO128, 128); };
      auto k_desc = [&]() { return make_desc(kring + (s_n % L_KSTAGES) * sm.k_bytes, VK * 16, 128); };
      // prologue: the scores of the first two virtual heads
      for (uint32_t i = 0; i < 2 && i < n_heads; ++i) {
        const uint32_t st = s_n % L_KSTAGES;
        if (s_v == 0) mbar_wait_relaxed(&qbf_full[s_t & 1], (s_t >> 1) & 1, 60);
        mbar_wait(&k_full[st], (s_n / L_KSTAGES) & 1);
        tc_fence_after();
        umma(tmem_base + LC_S + VK * (s_n & 1), q_desc(), k_desc(), idesc_s, 0);
        umma_commit(&s_full[s_n & 1]);
        umma_commit(&k_empty[st]);
        s_advance();
      }
      for (uint32_t n = 0; n < n_heads; ++n) {
        const uint32_t team = n & 1, use = n >> 1;
        const uint32_t stv = n % L_VSTAGES, stk = s_n % L_KSTAGES;
        const uint32_t more = n + 2 < n_heads ? 1u : 0u;
        LSTAMP(n - 2u * NV == 10u || n - 2u * NV == 11u, 642 + (n - 2u * NV - 10u) * 4);
        // all the barriers of this step in one round trip: lane 0 P_n complete, 1 the team's previous O read out,
        // 2 V image landed, 3 K image of S_{n+2} landed, 4 Q tiles of its tile written (first head of a tile)
        {
          uint32_t wb = bar(&p_full[team]), wp = use & 1;
          if (lane == 1) wb = bar(&o_empty[team]), wp = (use & 1) ^ 1;
          if (lane == 2) wb = bar(&v_full[stv]), wp = (n / L_VSTAGES) & 1;
          if (lane == 3 && more) wb = bar(&k_full[stk]), wp = (s_n / L_KSTAGES) & 1;
          if (lane == 4 && more && s_v == 0) wb = bar(&qbf_full[s_t & 1]), wp = (s_t >> 1) & 1;
#ifdef FFD_LAYER_STAMPS  // which barrier is the late one? (serial waits, one stamp each)
          if (n - 2u * NV == 10u || n - 2u * NV == 11u) {
            const uint32_t sl = 900 + (n - 2u * NV - 10u) * 8;
            LSTAMP(true, sl);
            mbar_wait(&p_full[team], use & 1);
            LSTAMP(true, sl + 1);
            mbar_wait(&o_empty[team], (use & 1) ^ 1);
            LSTAMP(true, sl + 2);
            mbar_wait(&v_full[stv], (n / L_VSTAGES) & 1);
            LSTAMP(true, sl + 3);
            if (more) mbar_wait(&k_full[stk], (s_n / L_KSTAGES) & 1);
            LSTAMP(true, sl + 4);
          }
#endif
          mbar_wait_lanes(wb, wp);
        }
        LSTAMP(n - 2u * NV < (uint32_t)NV, 500 + (n - 2u * NV) * 2);
        tc_fence_after();
        LSTAMP(n - 2u * NV == 10u || n - 2u * NV == 11u, 640 + (n - 2u * NV - 10u) * 4);
        // P.V of virtual head n and -- behind it on the in-order tensor pipe -- the scores of virtual head n + 2 into
        // the buffer P_n is read from
        attn_issue_step(tmem_base + LC_O + 32 * team, tmem_base + LC_S + VK * team,
                        make_desc(vring + stv * sm.v_bytes, LHP * 16, 128), VSTEP, IDESC_PV, bar(&o_full[team]),
                        bar(&v_empty[stv]), more, tmem_base + LC_S + VK * team, q_desc(), k_desc(), idesc_s,
                        bar(&s_full[team]), bar(&k_empty[stk]));
        LSTAMP(n - 2u * NV == 10u || n - 2u * NV == 11u, 641 + (n - 2u * NV - 10u) * 4);
        if (more) s_advance();
        LSTAMP(n - 2u * NV < (uint32_t)NV, 501 + (n - 2u * NV) * 2);
      }
      // drain: nothing of this warp is in flight when the CTA exits
      mbar_wait_relaxed(&o_full[(n_heads - 1) & 1], ((n_heads - 1) >> 1) & 1, 200);
    } else {
      // ================= FFN MMA issuer =================
      constexpr uint32_t IDESC1 = make_idesc(TILE_M, FCH);
      constexpr uint32_t IDESC2 = make_idesc(TILE_M, DP);
      static_assert(DP / 16 == 5 && FCH / 16 == 4, "ffn_issue_step is written for 5 + 4 K-steps");
      const uint32_t bar0 = smem_u32(bars);
      auto bar = [&](const uint64_t* b) { return bar0 + (uint32_t)((b - bars) * 8); };
      const uint32_t w1_a = smem_u32(smem + sm.w1), w2_a = smem_u32(smem + sm.w2);
      uint32_t c1 = 0, c2 = 0;  // ring consumption counters (W1 / W2 half-slices)
      uint32_t nh = 0;          // half-slices whose GEMM1 has been issued
      uint32_t ng = 0;          // half-slices whose GEMM2 has been issued
      for (uint32_t t = 0; t < my_tiles; ++t) {
        // X (bf16) and y1 + b2 (the Y accumulator) of tile t are in TMEM; the first two W1 half-slices have landed
        {
          uint32_t wb = bar(x_full), wp = t & 1;
          if (lane == 1) wb = bar(&f1_full[c1 % L_WSTAGES]), wp = (c1 / L_WSTAGES) & 1;
          if (lane == 2 && n_slices > 1) wb = bar(&f1_full[(c1 + 1) % L_WSTAGES]), wp = ((c1 + 1) / L_WSTAGES) & 1;
          mbar_wait_lanes(wb, wp);
        }
        tc_fence_after();
        LSTAMP(t == 2, 299);
        for (int i = 0; i < 2 && i < n_slices; ++i) {
          const uint32_t st = c1 % L_WSTAGES, b = nh & 1;
          umma_ts_chain<DP / 16>(tmem_base + LC_H + FCH * b, tmem_base + LC_X, 8,
                                 make_desc(w1_a + st * sm.w1_bytes, FCH * 16, 128), (2 * FCH * 16) >> 4, IDESC1, 0);
          umma_commit(&h_full[b]);
          umma_commit(&f1_empty[st]);
          ++c1, ++nh;
        }
        for (int j = 0; j < n_slices; ++j, ++ng) {
          const uint32_t b = ng & 1, st2 = c2 % L_WSTAGES, st1 = c1 % L_WSTAGES;
          const uint32_t more = j + 2 < n_slices ? 1u : 0u, last = j + 1 == n_slices ? 1u : 0u;
          LSTAMP(t == 2 && (j == 8 || j == 9), 600 + (j - 8) * 10 + 0);
          // one round trip: lane 0 relu(H_b) of half-slice j packed in place, 1 W2_j landed, 2 W1_{j+2} landed
          {
            uint32_t wb = bar(&r_full[b]), wp = (ng >> 1) & 1;
            if (lane == 1) wb = bar(&f2_full[st2]), wp = (c2 / L_WSTAGES) & 1;
            if (lane == 2 && more) wb = bar(&f1_full[st1]), wp = (c1 / L_WSTAGES) & 1;
            mbar_wait_lanes(wb, wp);
          }
          LSTAMP(t == 2 && (j == 8 || j == 9), 600 + (j - 8) * 10 + 1);
          tc_fence_after();
          LSTAMP(t == 2 && (j == 8 || j == 9), 600 + (j - 8) * 10 + 2);
          // Y += relu(H_b) . W2_j^T (on top of y1 + b2 from the first half-slice on), then -- behind it, because it
          // overwrites the columns GEMM2 reads -- H_b = X . W1_{j+2}^T; all commits under the same election
          ffn_issue_step(tmem_base + LC_Y, tmem_base + LC_H + FCH * b, make_desc(w2_a + st2 * sm.w2_bytes, DP * 16, 128),
                         (2 * DP * 16) >> 4, IDESC2, bar(&f2_empty[st2]), bar(y_full), last, more,
                         tmem_base + LC_H + FCH * b, tmem_base + LC_X, make_desc(w1_a + st1 * sm.w1_bytes, FCH * 16, 128),
                         (2 * FCH * 16) >> 4, IDESC1, bar(&h_full[b]), bar(&f1_empty[st1]));
          LSTAMP(t == 2 && (j == 8 || j == 9), 600 + (j - 8) * 10 + 3);
          ++c2;
          if (more) ++c1, ++nh;
          LSTAMP(t == 2 && j < 64, 300 + j);
        }
      }
      mbar_wait_relaxed(y_full, (my_tiles - 1) & 1, 200);  // drain
    }
  } else if (warp >= W_READ) {
    // ================= O read-out warps: combine a head's partial results, normalise, write the out-proj operand =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
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
        // the row sum sits in slot 7 of every accumulator (hd <= 7): the low 8 columns hold everything
        float acc[8];
        float mrun = 0.f;
        for (int c = 0; c < NCH; ++c, ++n) {
          const uint32_t team = n & 1;
          mbar_wait_relaxed(&o_full[team], (n >> 1) & 1, 32);
          tc_fence_after();
          LSTAMP(warp == W_READ && n - 2u * NV < (uint32_t)NV, 760 + (n - 2u * NV) * 2);
          uint32_t r[16];
          FFD_TMEM_LD8(tmem_base + lane_base + LC_O + 32 * team, r);
          FFD_TMEM_LD8(tmem_base + lane_base + LC_O + 32 * team + 16, (r + 8));
          const float* mh = mq + ((n & 3) * 2) * TILE_M + row;
          const float q0 = mh[0], q1 = mh[TILE_M];
          float mm = fmaxf(q0, q1);
          if (c > 0) mm = fmaxf(mm, mrun);
          const float w0 = exp2_poly3(q0 - mm), w1 = exp2_poly3(q1 - mm);
          const float wp = c > 0 ? exp2_poly3(mrun - mm) : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = c > 0 ? acc[j] * wp : 0.f;
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&o_empty[team]);
          LSTAMP(warp == W_READ && n - 2u * NV < (uint32_t)NV, 761 + (n - 2u * NV) * 2);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            fma2(acc[j], acc[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), w0, acc[j], acc[j + 1]);
            fma2(acc[j], acc[j + 1], __uint_as_float(r[8 + j]), __uint_as_float(r[9 + j]), w1, acc[j], acc[j + 1]);
          }
          mrun = mm;
        }
        const float inv = rcp_fma(acc[7]);
        // the out-projection of the previous tile must have read its operand before the first overwrite
        if (h == 0 && t >= 1) mbar_wait_relaxed(obf_empty, (t - 1) & 1, 60);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          if (j < hd) {
            const int col = h * hd + j;
            *reinterpret_cast<op16*>(ob + row_off + (col >> 3) * (TILE_M * 16) + (col & 7) * 2) =
                to_op(acc[j] * inv);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(obf_full);
    }
  } else if (warp >= W_TILE) {
    // ================= tile warps =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
    const uint32_t scr = tmem_base + lane_base + LC_H;
    const bool lead = warp == W_TILE;  // issues the scratch MMAs and the residual-tile loads (whole warp converged)
    constexpr uint32_t LBO128 = 128 * 16;
    constexpr int n4 = d / 4;
    const uint32_t xo_a = smem_u32(smem + sm.xo), obf_a = smem_u32(smem + sm.obf);
    const uint32_t wqo_a = smem_u32(smem + sm.wqo);
    uint32_t wc = 0;  // loads of the Wq / Wo buffer consumed so far (lead warp)
    uint8_t* xo = smem + sm.xo;
    float* res = reinterpret_cast<float*>(smem + sm.res);
    float* stg = reinterpret_cast<float*>(smem + sm.stage + quad * (32 * 80));  // this warp's [32][20] floats
    uint32_t cs = 0;  // scratch passes so far
    // constant columns of the Q-projection operand (d, d+1 = 1: bias hi/lo; the rest of the padding = 0)
    {
      uint4 pk = make_uint4(pack_op(1.f, 1.f), 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(xo + row_off + (d >> 3) * (TILE_M * 16)) = pk;
    }
    // residual rows of local tile t -> shared memory (one bulk copy; a tile of 128 rows is contiguous in HBM)
    // pull the rows of local tile t into L2 well before stage_xo reads them (first touch: HBM latency otherwise)
    auto prefetch_tile = [&](uint32_t t) {
      if (lead && lane == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x + tile_row0(t) * d),
                     "r"((uint32_t)(tile_rows(t) * d * 4))
                     : "memory");
    };
    auto load_residual = [&](uint32_t t) {
      if (lead && lane == 0) {
        const uint32_t bytes = (uint32_t)(tile_rows(t) * d * 4);
        mbar_expect_tx(t_full, bytes);
        bulk_load(res, x + tile_row0(t) * d, bytes, t_full);
      }
    };
    // one scratch pass: C[128 x N] = A[128 x DP] . W[N x DP]^T into the H columns (the FFN is between tiles).  W is
    // the current content of the Wq / Wo buffer; once the MMA has read it the lead warp requests the next content
    // (`next`, nullptr = none): Wq right behind the out-projection, Wo -- a whole tile ahead -- behind the Q projection.
    auto scratch_pass = [&](uint32_t a_addr, int N, uint64_t* also_commit, const uint8_t* next, uint32_t next_bytes) {
      tc_fence_before();
      named_bar_sync(1, 128);  // operand writes of all four warps are complete (and the previous pass's reads)
      if (lead) {
        mbar_wait(w_full, wc & 1);
        ++wc;
        tc_fence_after();
        umma_ss_chain<DP / 16>(tmem_base + LC_H, make_desc(a_addr, LBO128, 128), (2 * LBO128) >> 4,
                               make_desc(wqo_a, N * 16, 128), (2 * N * 16) >> 4, make_idesc(TILE_M, N), 0);
        umma_commit(scr_full);
        if (also_commit) umma_commit(also_commit);
      }
      mbar_wait(scr_full, cs & 1);
      ++cs;
      tc_fence_after();
      if (lead && lane == 0 && next) {
        mbar_expect_tx(w_full, next_bytes);
        bulk_load(smem + sm.wqo, next, next_bytes, w_full);
      }
    };
    // X of local tile t, read from global memory: a warp's 32 rows are one contiguous block, lane l takes its 16-byte
    // chunks l, l + 32, ... (coalesced); two neighbouring chunks make one 16-byte bf16 unit of the operand
    auto stage_xo = [&](uint32_t t) {
      const int rows = tile_rows(t);
      const int wrows = rows - quad * 32 < 32 ? rows - quad * 32 : 32;  // valid rows of this warp (may be <= 0)
      const float4* src = reinterpret_cast<const float4*>(x + (tile_row0(t) + quad * 32) * d);
      // 18 chunks per lane, in three batches of six loads that are all in flight together (one L2 round trip per
      // batch; issued one by one the loop costs a round trip per chunk: measured 11 K cycles per tile)
      static_assert((32 * n4) % (32 * 6) == 0, "stage_xo batches");
#pragma unroll 1
      for (int q0 = lane; q0 < 32 * n4; q0 += 32 * 6) {
        float4 a[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int q = q0 + 32 * i;
          a[i] = (q / n4 < wrows) ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int q = q0 + 32 * i;
          const int r = q / n4, c4 = q % n4;
          float4 b4;
          b4.x = __shfl_down_sync(0xffffffffu, a[i].x, 1);
          b4.y = __shfl_down_sync(0xffffffffu, a[i].y, 1);
          b4.z = __shfl_down_sync(0xffffffffu, a[i].z, 1);
          b4.w = __shfl_down_sync(0xffffffffu, a[i].w, 1);
          if ((lane & 1) == 0) {
            const int trow = quad * 32 + r;
            uint4 pk;
            pk.x = pack_op(a[i].x, a[i].y);
            pk.y = pack_op(a[i].z, a[i].w);
            pk.z = pack_op(b4.x, b4.y);
            pk.w = pack_op(b4.z, b4.w);
            *reinterpret_cast<uint4*>(xo + (uint32_t)(trow >> 3) * 128 + (uint32_t)(trow & 7) * 16 +
                                      (c4 >> 1) * (TILE_M * 16)) = pk;
          }
        }
      }
      fence_proxy_async();
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
    // Q projection of local tile t (its bf16 X operand has been staged) -> bf16 Q tiles
    // `then_wo`: the next user of the weight buffer is an out-projection (else another Q projection: Wq stays)
    auto project_q = [&](uint32_t t, bool then_wo) {
      uint8_t* qb = smem + sm.qbf + (t & 1) * sm.qbf_bytes;
      scratch_pass(xo_a, QN, nullptr, then_wo ? woimg : wqimg, then_wo ? sm.wo_bytes : sm.wq_bytes);
      for (int u = 0; u < QN; u += 32) {
        uint32_t ra[16], rb[16];
        const bool two = u + 16 < QN;
        FFD_TMEM_LD16(scr + u, ra);
        if (two) FFD_TMEM_LD16(scr + u + 16, rb);
        tmem_ld_wait();
        store_q16(ra, qb + row_off + (u >> 3) * (TILE_M * 16));
        if (two) store_q16(rb, qb + row_off + ((u >> 3) + 2) * (TILE_M * 16));
      }
      fence_proxy_async();
      mbar_arrive(&qbf_full[t & 1]);
    };
    // F1: out-projection of local tile t + residual -> LayerNorm1 -> y1: bf16 -> TMEM X, y1 + b2 (fp32) -> TMEM Y
    // `next_w`: what the weight buffer holds next -- 1 Wq (a Q projection follows), 2 Wo (the next user is the next
    // out-projection), 0 nothing (last pass of the kernel: no copy may be in flight when the CTA exits)
    auto finish_attention = [&](uint32_t t, int next_w) {
      float4* myrow = reinterpret_cast<float4*>(res) + (size_t)row * n4;
      float s1 = 0.f, s2 = 0.f;
      if (lead) mbar_wait_relaxed(obf_full, t & 1, 100);
      LSTAMP(warp == W_TILE && t == 2, 207);
      scratch_pass(obf_a, DP, obf_empty, next_w == 1 ? wqimg : next_w == 2 ? woimg : nullptr,
                   next_w == 1 ? sm.wq_bytes : sm.wo_bytes);
      mbar_wait_relaxed(t_full, t & 1, 100);  // the residual rows have landed
      LSTAMP(warp == W_TILE && t == 2, 208);
#pragma unroll
      for (int u = 0; u < DP; u += 16) {
        uint32_t r[16];
        FFD_TMEM_LD16(scr + u, r);
        float4 e[4];  // residual: issued before the TMEM round trip completes
#pragma unroll
        for (int j = 0; j < 4; ++j) e[j] = u + 4 * j < d ? myrow[(u >> 2) + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (u + 4 * j < d) {
            float4 v = e[j];
            v.x += __uint_as_float(r[4 * j]);
            v.y += __uint_as_float(r[4 * j + 1]);
            v.z += __uint_as_float(r[4 * j + 2]);
            v.w += __uint_as_float(r[4 * j + 3]);
            myrow[(u >> 2) + j] = v;
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
          }
        }
      }
      const float mean = s1 / (float)d;
      const float rstd = rsqrtf(fmaxf(s2 / (float)d - mean * mean, 0.f) + eps);
      const uint32_t xcol = tmem_base + lane_base + LC_X, ycol = tmem_base + lane_base + LC_Y;
#pragma unroll
      for (int u = 0; u < DP; u += 16) {
        uint32_t yv[16], xp[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = u + 4 * j;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 ob2 = o;
          if (c < d) {
            const float4 v = myrow[c >> 2];
            o.x = (v.x - mean) * rstd * consts[c] + consts[DP + c];
            o.y = (v.y - mean) * rstd * consts[c + 1] + consts[DP + c + 1];
            o.z = (v.z - mean) * rstd * consts[c + 2] + consts[DP + c + 2];
            o.w = (v.w - mean) * rstd * consts[c + 3] + consts[DP + c + 3];
            ob2.x = o.x + consts[2 * DP + c];
            ob2.y = o.y + consts[2 * DP + c + 1];
            ob2.z = o.z + consts[2 * DP + c + 2];
            ob2.w = o.w + consts[2 * DP + c + 3];
          } else if (c == d) {
            o.x = 1.f, o.y = 1.f;  // the constant-1 columns that carry b1 (hi + lo parts) through GEMM1
          }
          xp[2 * j] = pack_op(o.x, o.y);
          xp[2 * j + 1] = pack_op(o.z, o.w);
          yv[4 * j] = __float_as_uint(ob2.x);
          yv[4 * j + 1] = __float_as_uint(ob2.y);
          yv[4 * j + 2] = __float_as_uint(ob2.z);
          yv[4 * j + 3] = __float_as_uint(ob2.w);
        }
        FFD_TMEM_ST8(xcol + (u >> 1), xp);
        FFD_TMEM_ST16(ycol + u, yv);
      }
      tmem_st_wait();
      fence_proxy_async();  // the next residual tile (async proxy) is ordered after these reads of the buffer
    };
    // F2: Y (= y1 + b2 + FFN) -> LayerNorm2 -> global memory, 16 columns at a time through a small per-warp buffer so
    // that every store instruction writes whole 64-byte row pieces
    auto finish_ffn = [&](uint32_t t) {
      const int rows = tile_rows(t);
      const int wrows = rows - quad * 32 < 32 ? rows - quad * 32 : 32;
      const uint32_t ycol = tmem_base + lane_base + LC_Y;
      mbar_wait_relaxed(y_full, t & 1, 40);
      tc_fence_after();
      LSTAMP(warp == W_TILE && t == 2, 205);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int u = 0; u < DP; u += 16) {
        uint32_t r[16];
        FFD_TMEM_LD16(ycol + u, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (u + j < d) {
            const float v = __uint_as_float(r[j]);
            s1 += v;
            s2 = fmaf(v, v, s2);
          }
        }
      }
      const float mean = s1 / (float)d;
      const float rstd = rsqrtf(fmaxf(s2 / (float)d - mean * mean, 0.f) + eps);
      float* dst = y + (tile_row0(t) + quad * 32) * d;
#pragma unroll
      for (int u = 0; u < DP; u += 16) {
        if (u >= d) break;
        uint32_t r[16];
        FFD_TMEM_LD16(ycol + u, r);
        tmem_ld_wait();
        __syncwarp();  // the previous unit's reads of the staging buffer are complete
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = u + 4 * j;
          if (c < d) {
            float4 o;
            o.x = (__uint_as_float(r[4 * j]) - mean) * rstd * consts[3 * DP + c] + consts[4 * DP + c];
            o.y = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * consts[3 * DP + c + 1] + consts[4 * DP + c + 1];
            o.z = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * consts[3 * DP + c + 2] + consts[4 * DP + c + 2];
            o.w = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * consts[3 * DP + c + 3] + consts[4 * DP + c + 3];
            if (eps < 0.f) o = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            *reinterpret_cast<float4*>(stg + lane * 20 + 4 * j) = o;
          }
        }
        __syncwarp();
        // lane l stores the 16-byte pieces l, l + 32, ... of the [32 rows x 4 pieces] block
        const int pieces = (d - u) >= 16 ? 4 : (d - u) / 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = lane + 32 * i;
          const int rr = idx >> 2, pc = idx & 3;
          if (rr < wrows && pc < pieces)
            *reinterpret_cast<float4*>(dst + (size_t)rr * d + u + 4 * pc) =
                *reinterpret_cast<const float4*>(stg + rr * 20 + 4 * pc);
        }
      }
    };

    // ---- prologue: Q tiles of tile 0 (and 1), the bf16 X operand of tile 1, the residual rows of tile 0 ----
    load_residual(0);
    stage_xo(0);
    project_q(0, my_tiles == 1);
    if (my_tiles > 1) {
      stage_xo(1);
      project_q(1, true);
      if (my_tiles > 2) stage_xo(2);
    }
    // ---- period p >= 1: the heads of tile p are running (p = my_tiles: nothing is) ----
    for (uint32_t p = 1; p <= my_tiles; ++p) {
      const bool more_q = p + 1 < my_tiles;
      if (p + 2 < my_tiles) prefetch_tile(p + 2);
      LSTAMP(warp == W_TILE && p == 3, 200);
      finish_attention(p - 1, more_q ? 1 : p < my_tiles ? 2 : 0);
      LSTAMP(warp == W_TILE && p == 3, 201);
      if (more_q) project_q(p + 1, true);
      LSTAMP(warp == W_TILE && p == 3, 202);
      tc_fence_before();
      mbar_arrive(x_full);     // the FFN of tile p-1 may start (X, Y in TMEM; the scratch columns are free again)
      named_bar_sync(1, 128);  // all four warps are done with the residual buffer
      if (p < my_tiles) load_residual(p);
      LSTAMP(warp == W_TILE && p == 3, 203);
      if (p + 2 < my_tiles) stage_xo(p + 2);  // while the FFN of tile p-1 runs
      LSTAMP(warp == W_TILE && p == 3, 204);
      finish_ffn(p - 1);
      LSTAMP(warp == W_TILE && p == 3, 206);
    }
  } else if (warp >= W_EPI) {
    // ================= FFN epilogue warps: H_b (fp32) -> relu -> bf16 pairs, in place =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t total = my_tiles * (uint32_t)n_slices;
    for (uint32_t i = 0; i < total; ++i) {
      const uint32_t b = i & 1;
      const uint32_t hcol = tmem_base + lane_base + LC_H + FCH * b;
      mbar_wait_relaxed(&h_full[b], (i >> 1) & 1, 20);
      tc_fence_after();
      LSTAMP(warp == W_EPI && i - 2u * n_slices < 32u, 400 + (i - 2u * n_slices) * 2);
      uint32_t r[32], pk[16];
      FFD_TMEM_LD16(hcol, r);
      FFD_TMEM_LD16(hcol + 16, (r + 16));
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = relu_pack_op(r[2 * e], r[2 * e + 1]);
      FFD_TMEM_LD16(hcol + 32, r);
      FFD_TMEM_LD16(hcol + 48, (r + 16));
      tmem_ld_wait();
      FFD_TMEM_ST16(hcol, pk);  // columns 0..15 <- hidden units 0..31 (their fp32 values are in registers / done)
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = relu_pack_op(r[2 * e], r[2 * e + 1]);
      FFD_TMEM_ST16(hcol + 16, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&r_full[b]);
      LSTAMP(warp == W_EPI && i - 2u * n_slices < 32u, 401 + (i - 2u * n_slices) * 2);
    }
  } else {
    // ================= softmax teams =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
    const uint32_t team = (uint32_t)(warp - W_SOFTMAX) >> 3;
    const int grp = ((warp - W_SOFTMAX) >> 2) & 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t scol = tmem_base + lane_base + LC_S + VK * team + 48 * grp;  // this warp's 48 score columns
    for (uint32_t n = team; n < n_heads; n += 2) {
      mbar_wait(&s_full[team], (n >> 1) & 1);
      tc_fence_after();
      LSTAMP((warp == 0 || warp == 8) && n - 2u * NV < (uint32_t)NV, (n - 2u * NV) * 2);
      uint32_t r[48];
      FFD_TMEM_LD16(scol, r);
      FFD_TMEM_LD16(scol + 16, (r + 16));
      FFD_TMEM_LD16(scol + 32, (r + 32));
      tmem_ld_wait();
      // row maximum over this group's 48 keys (three independent chains).  Padded keys need no mask: slot 15 of
      // every Q tile is the constant 1 and slot 15 of a padded K row is -29952 (weight 0 in the combine).
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        m0 = max3(m0, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
        m1 = max3(m1, __uint_as_float(r[16 + e]), __uint_as_float(r[17 + e]));
        m2 = max3(m2, __uint_as_float(r[32 + e]), __uint_as_float(r[33 + e]));
      }
      const float mb = max3(m0, m1, m2);
      mq[((n & 3) * 2 + grp) * TILE_M + row] = mb;
      const float mc = -mb;
      // p = exp2(s - m) as fp16 pairs, written over this warp's own score columns: 48 fp32 -> 24 packed columns
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2)
#ifdef FFD_L_NO_EXP
          pk[e >> 1] = r[16 * j + e] + r[16 * j + e + 1] + __float_as_uint(mc);
#else
          pk[e >> 1] = ex2_pack_f16_shift(__uint_as_float(r[16 * j + e]), __uint_as_float(r[16 * j + e + 1]), mc);
#endif
        FFD_TMEM_ST8(scol + 8 * j, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[team]);
      LSTAMP((warp == 0 || warp == 8) && n - 2u * NV < (uint32_t)NV, (n - 2u * NV) * 2 + 1);
      LSTAMP((warp & 3) == 1 && n - 2u * NV < (uint32_t)NV, 700 + (n - 2u * NV) * 2 + grp);  // quad 1, both groups
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_AMMA) tmem_dealloc(tmem_base, 512);
#undef LSTAMP
}

}  // namespace tc

// ---- host side ------------------------------------------------------------------------------------------------
extern long long* g_attn_dbg;  // optional device buffer of clock64 stamps (ffd_debug_set_attn_stamps)

static int layer_key_chunks(int L) { return (L + tc::VK - 1) / tc::VK; }

bool tc_layer_supported(const ffd_dims& m) {
  if (!tc_attn_supported(m) || m.d_model != 72) return false;
  const int hd = m.d_model / m.n_head;
  if (hd > 7 || (m.d_ff % tc::FCH) != 0 || m.d_ff < 2 * tc::FCH) return false;
  const int G = 15 / hd < m.n_head ? 15 / hd : m.n_head;
  const int QN = (m.n_head + G - 1) / G * tc::LHP;
  if (QN > 128) return false;  // scratch pass of the Q projection: one MMA into the 128 H columns
  return tc::layer_smem(80, QN, 72).total <= 232448;
}

int64_t tc_kv96_image_bytes(const ffd_dims& m) {
  return (int64_t)m.n_layers * m.n_head * layer_key_chunks(m.max_len) * 2 * tc::VK * tc::LHP * 2;
}

// OPT-IN (FFD_FUSED_LAYER=1 / ffd_debug_set_fused_layer(1)); the default is the two-kernel layer.  Measured on B200 at
// the ECG shapes (B = 16384): 5.8 ms per layer against 1.98 + 1.34 ms for the attention and FFN kernels.  The kernel
// is correct (tests/test_gpu_layer_kernel.py) and both pipes do run side by side, but with ONE tile per pipeline stage
// every step is a chain of mbarrier / commit hand-offs of 150-300 cycles each: 32 FFN half-slice steps of ~850
// cycles (535 in an otherwise idle SM, tools/ffn_pipe_bench.cu; 320 of them tensor time), 24 virtual-head steps of
// ~1100 + turn-around, and the tile warps' serial chain around them.  With MMAs and exponentials switched off the
// hand-off skeleton alone takes 4.7 ms (DESIGN.md section 8 has the timeline).
static int g_fused_layer_mode = -1;  // -1 unknown (FFD_FUSED_LAYER or the default), 0 off, 1 on
void tc_debug_set_fused_layer(int mode) { g_fused_layer_mode = mode; }
bool tc_layer_enabled(const ffd_dims& m) {
  int mode = g_fused_layer_mode;
  if (mode < 0) {
    static const int env_mode = [] {
      const char* e = getenv("FFD_FUSED_LAYER");
      return e && e[0] == '1' ? 1 : 0;
    }();
    mode = env_mode;
  }
  return mode == 1 && tc_layer_supported(m);
}

// y = LN2(y1 + FFN(y1)), y1 = LN1(x + cached attention(x)) for layer `layer`, rows [0, M); y may alias x.
// kvimg96: tc_pack_kv96 images of ALL layers.
int tc_layer_cached(const ffd_dims& m, const void* packed, const void* kvimg96, const float* weights, int layer,
                    const float* x, float* y, int64_t M, cudaStream_t s) {
  constexpr int DP = 80, D = 72;
  const Arena ar = arena_of(m);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const int hd = m.d_model / m.n_head;
  const int G = 15 / hd < m.n_head ? 15 / hd : m.n_head;
  const int QN = (m.n_head + G - 1) / G * tc::LHP;
  const int NCH = layer_key_chunks(m.max_len);
  const tc::LayerSmem sm = tc::layer_smem(DP, QN, D);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    FFD_CHECK_CUDA(cudaGetDevice(&dev));
    FFD_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    FFD_CHECK_CUDA(cudaFuncSetAttribute(tc::layer_cached_tc_kernel<DP, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        sm.total));
  }
  const uint8_t* attn_img = (const uint8_t*)tc_attn_image(m, packed, layer);
  const uint8_t* ffn_img = (const uint8_t*)tc_ffn_image(m, packed, layer);
  const uint8_t* w1h = (const uint8_t*)tc_ffn_half_image(m, packed, layer);  // W1 as [64 n x DP k] half-slices
  const int n128 = m.d_ff / 128;
  const uint8_t* w2 = ffn_img + (size_t)n128 * 128 * DP * 2;                  // W2 [DP n x 128 k] = two K-halves each
  const uint8_t* kvl = (const uint8_t*)kvimg96 + (size_t)layer * m.n_head * NCH * 2 * tc::VK * tc::LHP * 2;
  const int64_t n_tiles = (M + tc::TILE_M - 1) / tc::TILE_M;
  const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
  static const float eps = getenv("FFD_FFN_RAW") ? -1.f : 1e-5f;
  tc::layer_cached_tc_kernel<DP, D><<<grid, tc::LAYER_THREADS, sm.total, s>>>(
      x, y, attn_img, attn_img + (size_t)QN * DP * 2, kvl, w1h, w2, w + ar.n1w, w + ar.n1b, w + ar.b2, w + ar.n2w,
      w + ar.n2b, M, m.n_head, hd, G, QN, NCH, m.d_ff / tc::FCH, n_tiles, eps, g_attn_dbg);
  FFD_CHECK_LAUNCH("layer_cached_tc_kernel");
  return FFD_OK;
}

}  // namespace ffd
