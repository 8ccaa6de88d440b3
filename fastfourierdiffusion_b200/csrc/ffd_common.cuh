// Shared helpers for the ffd_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ffd_b200.h"

namespace ffd {

void set_error(const char* fmt, ...);
void count_launch();

#define FFD_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ffd::set_error(__VA_ARGS__);        \
      return FFD_ERR_ARG;                 \
    }                                     \
  } while (0)

#define FFD_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ffd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FFD_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define FFD_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    ffd::count_launch();                                                                  \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ffd::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
      return FFD_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---- weight arena offsets (ffd_b200.h) ---------------------------------------
struct Arena {
  int64_t pos, emb_w, emb_b, unemb_w, unemb_b, layer0, layer_stride, total;
  // within a layer block
  int64_t in_w, in_b, out_w, out_b, w1, b1, w2, b2, n1w, n1b, n2w, n2b;
};

static inline Arena arena_of(const ffd_dims& m) {
  Arena a;
  const int64_t L = m.max_len, C = m.n_channels, d = m.d_model, F = m.d_ff;
  int64_t o = 0;
  a.pos = o; o += L * d;
  a.emb_w = o; o += d * C;
  a.emb_b = o; o += d;
  a.unemb_w = o; o += C * d;
  a.unemb_b = o; o += C;
  a.layer0 = o;
  int64_t q = 0;
  a.in_w = q; q += 3 * d * d;
  a.in_b = q; q += 3 * d;
  a.out_w = q; q += d * d;
  a.out_b = q; q += d;
  a.w1 = q; q += F * d;
  a.b1 = q; q += F;
  a.w2 = q; q += d * F;
  a.b2 = q; q += d;
  a.n1w = q; q += d;
  a.n1b = q; q += d;
  a.n2w = q; q += d;
  a.n2b = q; q += d;
  a.layer_stride = q;
  a.total = o + q * m.n_layers;
  return a;
}

// ---- key chunking of the tcgen05 attention kernel (ffd_tc_attn.cu) ---------------
// The cached keys of a head are processed in NCH chunks of CK keys (CK a multiple of 64, at most 192).
static inline int attn_key_chunks(int L) { return (L + 191) / 192; }
static inline int attn_chunk_keys(int L) {
  const int n = attn_key_chunks(L);
  return ((L + n - 1) / n + 63) / 64 * 64;
}

// ---- workspace carve-up --------------------------------------------------------
struct Workspace {
  float* h;     // [M, d] residual stream
  float* tmp;   // [M, d] pre-LN sums
  float* attn;  // [M, d] attention output (heads merged)
  float* qkv;   // [M, 3d]
  float* kvtmp; // [L, 2d] sample-0 K/V of a full step
  void* kvimg;  // bf16 K/V operand images of the cache for the tcgen05 attention kernel
  void* kvimg_own;  // the same per sample (stock attention on the tcgen05 kernel)
  void* kvimg96;    // the cache images again in chunks of 96 keys (fused layer kernel)
  int64_t bytes;
};

// Samples per pass of the per-sample K/V operand images (stock / mixed attention on the tcgen05 kernel): the images are
// 147 KB per ECG sample, so a full step walks the batch in groups of this many samples instead of holding B images.
constexpr int64_t OWN_CHUNK = 2048;

// lean: the tensor-core path covers every branch of the model (bf16 precision, tcgen05 attention with the fused K/V
// projection), so only the residual stream and the operand images are needed -- 54 KB per ECG sample instead of
// 323 KB + 147 KB, which is what lets one sampler batch hold 2^20 samples in 180 GB.
static inline Workspace carve(const ffd_dims& m, int64_t B, void* base, bool lean) {
  Workspace w;
  const int64_t M = B * m.max_len, d = m.d_model;
  char* p = (char*)base;
  int64_t o = 0;
  auto take = [&](int64_t nfloat) {
    float* r = (float*)(p + o);
    o += align_up(nfloat * 4, 1024);
    return r;
  };
  w.h = take(M * d);
  w.tmp = lean ? nullptr : take(M * d);
  w.attn = lean ? nullptr : take(M * d);
  w.qkv = lean ? nullptr : take(M * 3 * d);
  w.kvtmp = take((int64_t)m.max_len * 2 * d);
  // [NL, H, NCH] x (K image [CK x 16] bf16 + V image [16 x CK] fp16)
  w.kvimg = take((int64_t)m.n_layers * m.n_head * attn_key_chunks(m.max_len) * attn_chunk_keys(m.max_len) * 16);
  // [min(B, OWN_CHUNK), H, NCH] x (K image + V image): 64 bytes per key slot
  const int64_t own = B < OWN_CHUNK ? B : OWN_CHUNK;
  w.kvimg_own = take(own * m.n_head * attn_key_chunks(m.max_len) * attn_chunk_keys(m.max_len) * 16);
  // [NL, H, ceil(L / 96)] x (K image [96 x 16] bf16 + V image [16 x 96] fp16)
  w.kvimg96 = take((int64_t)m.n_layers * m.n_head * ((m.max_len + 95) / 96) * 96 * 16);
  w.bytes = o;
  return w;
}

// ---- per-step parameters read from DEVICE memory (CUDA-graph replay of the step loop, ffd_api.cu) ----------------
// A graph node's arguments are frozen at capture; what changes from step to step (time-embedding row, SDE scalars,
// Philox stream, injected-noise slice) is looked up through a step counter that a one-thread kernel advances.
struct StepCursor {
  int32_t ctr;          // absolute index of the step being executed
  int32_t noise_step0;  // absolute step index of noise[0]
  uint64_t seed;
  uint64_t stream_base; // Philox stream id of step 0
};
struct StepSource {     // all DEVICE pointers; cur == nullptr: arguments are taken by value as usual
  const StepCursor* cur;
  const ffd_step_coef* coefs;  // [n_steps]
};

// ---- device helpers --------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox-4x32-10 (Salmon et al. 2011); restated in oracle/fdiff_oracle.py:philox4x32_10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// four N(0,1) draws for element quad q of stream `stream_id`
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t stream_id, uint64_t q) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)stream_id,
                                           (uint32_t)(stream_id >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
  const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
  const float rad0 = sqrtf(-2.0f * logf(u0));
  const float rad1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  return make_float4(rad0 * c0, rad0 * s0, rad1 * c1, rad1 * s1);
}

}  // namespace ffd
