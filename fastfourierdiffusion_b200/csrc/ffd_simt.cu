// fp32 CUDA-core kernels of the sampling path: the FFD_PREC_FP32 arithmetic and every
// non-GEMM op (embedding, attention, LayerNorm, SDE update, noise, boundary DFT).
// Reference call sites are cited per kernel; paths are relative to the upstream repo.
#include "ffd_internal.h"

namespace ffd {

// ------------------------------------------------------------------------------------
// embed: h[m,:] = X[m,:] @ We^T + be + pos[l,:] + te[:]
//   score_models.py:96-102, transformer.py:26-28 (pos), :80-91 (time; one t per batch)
// ------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ X, const float* __restrict__ We,
                             const float* __restrict__ be, const float* __restrict__ pos,
                             const float* __restrict__ te, float* __restrict__ h, int64_t M, int L, int C,
                             int d, const int32_t* __restrict__ step_ctr) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * d) return;
  if (step_ctr) te += (int64_t)*step_ctr * d;
  const int64_t m = idx / d;
  const int j = (int)(idx - m * d);
  const int l = (int)(m % L);
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(X[m * C + c], We[j * C + c], acc);
  // same association as the reference: ((x@W + b) + pos) + te
  h[idx] = ((acc + be[j]) + pos[(int64_t)l * d + j]) + te[j];
}

// float4 variant (d % 4 == 0): a pure HBM write.  One block writes EMB_ROWS consecutive token rows (a contiguous span
// of the output) as 16-byte stores, several per thread in flight, with 32-bit index arithmetic only.
constexpr int EMB_ROWS = 128;
__global__ void __launch_bounds__(256)
embed4_kernel(const float* __restrict__ X, const float* __restrict__ We, const float* __restrict__ be,
              const float* __restrict__ pos, const float* __restrict__ te, float4* __restrict__ h, int64_t M, int L,
              int C, int d4, const int32_t* __restrict__ step_ctr) {
  if (step_ctr) te += (int64_t)*step_ctr * (d4 * 4);
  const int64_t m0 = (int64_t)blockIdx.x * EMB_ROWS;
  const int rows = M - m0 < EMB_ROWS ? (int)(M - m0) : EMB_ROWS;
  const int l0 = (int)(m0 % L);
  const int n = rows * d4;
  float4* out = h + m0 * d4;
  const float* x0 = X + m0 * C;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += 256) {
    const int r = i / d4;
    const int j = (i - r * d4) * 4;
    int l = l0 + r;
    l -= l >= L ? L : 0;
    l -= l >= L ? L : 0;  // EMB_ROWS may exceed L
    while (l >= L) l -= L;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) {
      const float xv = __ldg(x0 + r * C + c);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = fmaf(xv, __ldg(We + (j + e) * C + c), acc[e]);
    }
    const float4 b4 = *reinterpret_cast<const float4*>(be + j);
    const float4 p4 = *reinterpret_cast<const float4*>(pos + l * (d4 * 4) + j);
    const float4 t4 = *reinterpret_cast<const float4*>(te + j);
    float4 o;
    o.x = ((acc[0] + b4.x) + p4.x) + t4.x;
    o.y = ((acc[1] + b4.y) + p4.y) + t4.y;
    o.z = ((acc[2] + b4.z) + p4.z) + t4.z;
    o.w = ((acc[3] + b4.w) + p4.w) + t4.w;
    __stcs(out + i, o);  // streaming store: the first attention block reads it from HBM anyway (882 MB >> L2)
  }
}

int launch_embed(const float* X, const float* We, const float* be, const float* pos, const float* te, float* h,
                 int64_t M, int L, int C, int d, cudaStream_t s, const int32_t* step_ctr) {
  const bool aligned = d % 4 == 0 && ((uintptr_t)be % 16 == 0) && ((uintptr_t)pos % 16 == 0) &&
                       ((uintptr_t)te % 16 == 0) && ((uintptr_t)h % 16 == 0);
  if (aligned) {
    embed4_kernel<<<(unsigned)((M + EMB_ROWS - 1) / EMB_ROWS), 256, 0, s>>>(X, We, be, pos, te, (float4*)h, M, L, C,
                                                                              d / 4, step_ctr);
    FFD_CHECK_LAUNCH("embed4_kernel");
    return FFD_OK;
  }
  const int64_t n = M * d;
  embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(X, We, be, pos, te, h, M, L, C, d, step_ctr);
  FFD_CHECK_LAUNCH("embed_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// linear: Y[m,n] = sum_k A[m,k] W[n,k] + b[n]  (F.linear; cached_transformer.py:157-158,229,284-285)
// 64x64 tile, BK=16, 256 threads, 4x4 micro-tile.
// ------------------------------------------------------------------------------------
constexpr int LBM = 64, LBN = 64, LBK = 16;

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
              const float* __restrict__ bias, float* __restrict__ Y, int ldy, int64_t M, int N, int K,
              int relu) {
  __shared__ float As[LBK][LBM + 4];
  __shared__ float Ws[LBK][LBN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * LBM;
  const int n0 = blockIdx.y * LBN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += LBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;  // 0..1023 over a 64x16 tile, k fastest
      const int r = e >> 4, kk = e & 15;
      const int64_t m = m0 + r;
      const int k = k0 + kk;
      As[kk][r] = (m < M && k < K) ? A[m * lda + k] : 0.f;
      const int n = n0 + r;
      Ws[kk][r] = (n < N && k < K) ? W[(int64_t)n * ldw + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      Y[m * ldy + n] = v;
    }
  }
}

int launch_linear(const float* A, int lda, const float* W, int ldw, const float* bias, float* Y, int ldy,
                  int64_t M, int N, int K, int relu, cudaStream_t s) {
  if (M == 0 || N == 0) return FFD_OK;
  dim3 grid((unsigned)((M + LBM - 1) / LBM), (unsigned)((N + LBN - 1) / LBN));
  linear_kernel<<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, Y, ldy, M, N, K, relu);
  FFD_CHECK_LAUNCH("linear_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// attention: one CTA per (sample, head); a thread owns one query row and streams the keys
// with an online softmax.  K/V row j comes from the sample's own projection (qkv buffer)
// when j is in the recompute set, else from the sample-0 cache -- this one kernel covers
// nn.MultiheadAttention (all own), pure cache (none own) and the mixed branch
// (cached_transformer.py:125,237-305,309-311).
// ------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ qkv, int ldq, const float* __restrict__ kc,
                 const float* __restrict__ vc, const int32_t* __restrict__ own_mask, int n_prefix_own,
                 float* __restrict__ out, int L, int d, int H, int hd, float scale) {
  extern __shared__ float sm[];
  float* Ks = sm;                 // [L][HDP]
  float* Vs = sm + (size_t)L * HDP;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int64_t row0 = (int64_t)b * L;
  for (int e = threadIdx.x; e < L * HDP; e += blockDim.x) {
    const int j = e / HDP, c = e % HDP;
    float kv = 0.f, vv = 0.f;
    if (c < hd) {
      const bool own = own_mask ? (own_mask[j] != 0) : (j < n_prefix_own);
      if (own) {
        kv = qkv[(row0 + j) * ldq + d + h * hd + c];
        vv = qkv[(row0 + j) * ldq + 2 * d + h * hd + c];
      } else {
        kv = kc[((int64_t)h * L + j) * hd + c];
        vv = vc[((int64_t)h * L + j) * hd + c];
      }
    }
    Ks[e] = kv;
    Vs[e] = vv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float q[HDP], acc[HDP];
#pragma unroll
    for (int c = 0; c < HDP; ++c) {
      q[c] = (c < hd) ? qkv[(row0 + i) * ldq + h * hd + c] : 0.f;
      acc[c] = 0.f;
    }
    float mx = -INFINITY, sum = 0.f;
    for (int j = 0; j < L; ++j) {
      float sc = 0.f;
#pragma unroll
      for (int c = 0; c < HDP; ++c) sc = fmaf(q[c], Ks[j * HDP + c], sc);
      sc *= scale;
      if (sc > mx) {
        const float f = expf(mx - sc);
        sum *= f;
#pragma unroll
        for (int c = 0; c < HDP; ++c) acc[c] *= f;
        mx = sc;
      }
      const float p = expf(sc - mx);
      sum += p;
#pragma unroll
      for (int c = 0; c < HDP; ++c) acc[c] = fmaf(p, Vs[j * HDP + c], acc[c]);
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < HDP; ++c)
      if (c < hd) out[(row0 + i) * d + h * hd + c] = acc[c] * inv;
  }
}

int launch_attention(const float* qkv, int ldq, const float* kc, const float* vc, const int32_t* own_mask,
                     int n_prefix_own, float* out, int64_t B, int L, int d, int H, cudaStream_t s) {
  const int hd = d / H;
  const float scale = 1.0f / sqrtf((float)hd);
  const unsigned grid = (unsigned)(B * H);
  const int threads = 128;
#define FFD_ATT(HDP)                                                                                  \
  {                                                                                                   \
    const size_t smem = (size_t)2 * L * HDP * sizeof(float);                                          \
    if (smem > 48 * 1024)                                                                             \
      FFD_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                \
    attention_kernel<HDP><<<grid, threads, smem, s>>>(qkv, ldq, kc, vc, own_mask, n_prefix_own, out, L, d, H, \
                                                      hd, scale);                                     \
  }
  if (hd <= 8) FFD_ATT(8)
  else if (hd <= 16) FFD_ATT(16)
  else if (hd <= 32) FFD_ATT(32)
  else if (hd <= 64) FFD_ATT(64)
  else {
    set_error("head_dim %d > 64 unsupported", hd);
    return FFD_ERR_UNSUPPORTED;
  }
#undef FFD_ATT
  FFD_CHECK_LAUNCH("attention_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// add + LayerNorm: y = LN(res + x) * w + b, eps 1e-5, one warp per row
//   cached_transformer.py:126-127,131-132,320-321,326-327
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
add_ln_kernel(const float* __restrict__ res, const float* __restrict__ x, const float* __restrict__ w,
              const float* __restrict__ b, float* __restrict__ y, int64_t M, int d, float eps) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float v[8];  // d <= 256
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    v[i] = (c < d) ? res[row * d + c] + x[row * d + c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    const float t = (c < d) ? v[i] - mean : 0.f;
    q = fmaf(t, t, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    if (c < d) y[row * d + c] = (v[i] - mean) * rstd * w[c] + b[c];
  }
}

int launch_add_ln(const float* res, const float* x, const float* w, const float* b, float* y, int64_t M, int d,
                  cudaStream_t s) {
  if (d > 256) {
    set_error("d_model %d > 256 unsupported", d);
    return FFD_ERR_UNSUPPORTED;
  }
  add_ln_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(res, x, w, b, y, M, d, 1e-5f);
  FFD_CHECK_LAUNCH("add_ln_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// fused FFN (fp32): out = W2 relu(W1 x + b1) + b2 for a 64-row tile; the F-wide hidden
// activation lives in shared memory 64 columns at a time and never reaches HBM.
//   cached_transformer.py:130,325 (linear2(relu(linear1(src))))
// ------------------------------------------------------------------------------------
constexpr int FBM = 64, FFC = 64;

__global__ void __launch_bounds__(256)
ffn_kernel(const float* __restrict__ X, const float* __restrict__ W1, const float* __restrict__ b1,
           const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ out, int64_t M,
           int d, int F) {
  extern __shared__ float sm[];
  const int ldx = d + 1;
  float* Xs = sm;                    // [64][d+1]
  float* W1s = Xs + FBM * ldx;       // [64 f][d+1]
  float* Hs = W1s + FFC * ldx;       // [64][65]
  float* W2s = Hs + FBM * (FFC + 1); // [d][65]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * FBM;
  for (int e = tid; e < FBM * d; e += 256) {
    const int r = e / d, k = e % d;
    Xs[r * ldx + k] = (m0 + r < M) ? X[(m0 + r) * d + k] : 0.f;
  }
  float acc[4][8];  // rows ty*4+i, cols tx + 16*j  (d <= 128)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int f0 = 0; f0 < F; f0 += FFC) {
    __syncthreads();  // previous chunk's readers of W1s/W2s/Hs are done (and Xs is visible)
    for (int e = tid; e < FFC * d; e += 256) {
      const int r = e / d, k = e % d;
      W1s[r * ldx + k] = (f0 + r < F) ? W1[(int64_t)(f0 + r) * d + k] : 0.f;
    }
    for (int e = tid; e < d * FFC; e += 256) {
      const int n = e / FFC, f = e % FFC;
      W2s[n * (FFC + 1) + f] = (f0 + f < F) ? W2[(int64_t)n * F + f0 + f] : 0.f;
    }
    __syncthreads();
    float hreg[4][4] = {};
    for (int k = 0; k < d; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Xs[(ty * 4 + i) * ldx + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = W1s[(tx + 16 * j) * ldx + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) hreg[i][j] = fmaf(a[i], w[j], hreg[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = tx + 16 * j;
        const float bb = (f0 + f < F) ? b1[f0 + f] : 0.f;
        Hs[(ty * 4 + i) * (FFC + 1) + f] = fmaxf(hreg[i][j] + bb, 0.f);
      }
    __syncthreads();
    for (int f = 0; f < FFC; ++f) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Hs[(ty * 4 + i) * (FFC + 1) + f];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = tx + 16 * j;
        if (n < d) {
          const float w = W2s[n * (FFC + 1) + f];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(a[i], w, acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = tx + 16 * j;
      if (n < d) out[m * d + n] = acc[i][j] + b2[n];
    }
  }
}

int launch_ffn(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, float* out,
               int64_t M, int d, int F, cudaStream_t s) {
  if (d > 128) {
    set_error("fp32 FFN kernel supports d_model <= 128 (got %d)", d);
    return FFD_ERR_UNSUPPORTED;
  }
  const size_t smem = sizeof(float) * ((size_t)FBM * (d + 1) + (size_t)FFC * (d + 1) + (size_t)FBM * (FFC + 1) +
                                       (size_t)d * (FFC + 1));
  FFD_CHECK_CUDA(cudaFuncSetAttribute(ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ffn_kernel<<<(unsigned)((M + FBM - 1) / FBM), 256, smem, s>>>(X, W1, b1, W2, b2, out, M, d, F);
  FFD_CHECK_LAUNCH("ffn_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// KV store: keep sample 0's K/V rows (caching.py:326-328,389-391), layout [H, L, hd].
// src row l holds K at column k_off + h*hd + c and V at v_off + ...
// ------------------------------------------------------------------------------------
__global__ void scatter_kv_kernel(const float* __restrict__ src, int lds, int k_off, int v_off,
                                  const int32_t* __restrict__ own_mask, int n_prefix_own,
                                  float* __restrict__ kc, float* __restrict__ vc, int L, int H, int hd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = H * hd;
  if (idx >= L * d) return;
  const int l = idx / d, j = idx % d, h = j / hd, c = j % hd;
  const bool own = own_mask ? (own_mask[l] != 0) : (l < n_prefix_own);
  if (!own) return;
  kc[((int64_t)h * L + l) * hd + c] = src[(int64_t)l * lds + k_off + j];
  vc[((int64_t)h * L + l) * hd + c] = src[(int64_t)l * lds + v_off + j];
}

int launch_scatter_kv(const float* src, int lds, int k_off, int v_off, const int32_t* own_mask, int n_prefix_own,
                      float* kc, float* vc, int L, int H, int hd, cudaStream_t s) {
  const int n = L * H * hd;
  scatter_kv_kernel<<<(n + 255) / 256, 256, 0, s>>>(src, lds, k_off, v_off, own_mask, n_prefix_own, kc, vc, L, H,
                                                    hd);
  FFD_CHECK_LAUNCH("scatter_kv_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// unembed + reverse-SDE Euler-Maruyama update, one warp per token row.
//   score_models.py:113; sde.py:228-244 (VP), :143-163 (VE).  The update keeps the
//   reference's rounding order (no FMA contraction): diff = coef*G; diff2 = diff*diff;
//   drift = c_x*x - diff2*s; out = (x - drift*dt) + sqrt_dt*(diff*z).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float sde_update(const ffd_step_coef& c, float g, float x, float sc, float z) {
  const float diff = __fmul_rn(c.coef, g);
  const float diff2 = __fmul_rn(diff, diff);
  float drift;
  if (c.kind == FFD_SDE_VP)
    drift = __fsub_rn(__fmul_rn(c.c_x, x), __fmul_rn(diff2, sc));
  else
    drift = -__fmul_rn(diff2, sc);
  const float a = __fsub_rn(x, __fmul_rn(drift, c.dt));
  return __fadd_rn(a, __fmul_rn(c.sqrt_dt, __fmul_rn(diff, z)));
}

__global__ void __launch_bounds__(256)
unembed_step_kernel(const float* __restrict__ h, const float* __restrict__ Wu, const float* __restrict__ bu,
                    float* __restrict__ score_out, int do_step, ffd_step_coef coef,
                    const float* __restrict__ G, const float* __restrict__ x, const float* __restrict__ z,
                    uint64_t seed, uint64_t stream_id, float* __restrict__ x_out, int64_t M, int L, int C,
                    int d, StepSource src) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  if (src.cur) {  // per-step parameters from device memory (captured step loop)
    const StepCursor sc = *src.cur;
    coef = src.coefs[sc.ctr];
    seed = sc.seed;
    stream_id = sc.stream_base + (uint64_t)sc.ctr;
    if (z) z += (int64_t)(sc.ctr - sc.noise_step0) * (M * C);
  }
  const int lane = threadIdx.x & 31;
  float hv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane + 32 * i;
    hv[i] = (k < d) ? h[row * d + k] : 0.f;
  }
  const float g = do_step ? G[row % L] : 0.f;
  for (int c = 0; c < C; ++c) {
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      if (k < d) p = fmaf(hv[i], Wu[c * d + k], p);
    }
    p = warp_sum(p) + bu[c];
    if (lane == 0) {
      const int64_t e = row * C + c;
      if (score_out) score_out[e] = p;
      if (do_step) {
        float zz;
        if (z) {
          zz = z[e];
        } else {
          const float4 n4 = philox_normal4(seed, stream_id, (uint64_t)e >> 2);
          const int r = (int)(e & 3);
          zz = r == 0 ? n4.x : r == 1 ? n4.y : r == 2 ? n4.z : n4.w;
        }
        x_out[e] = sde_update(coef, g, x[e], p, zz);
      }
    }
  }
}

// tiled variant: a block stages 128 rows of h through shared memory with coalesced 16-byte loads, then one
// thread per row forms the C dot products (row stride d+1 words: conflict-free) -- HBM-bound read of h.
constexpr int UT = 128;
__global__ void __launch_bounds__(UT)
unembed_step_tiled_kernel(const float* __restrict__ h, const float* __restrict__ Wu, const float* __restrict__ bu,
                          float* __restrict__ score_out, int do_step, ffd_step_coef coef,
                          const float* __restrict__ G, const float* __restrict__ x, const float* __restrict__ z,
                          uint64_t seed, uint64_t stream_id, float* __restrict__ x_out, int64_t M, int L, int C,
                          int d, StepSource step_src) {
  extern __shared__ float sm[];
  if (step_src.cur) {  // per-step parameters from device memory (captured step loop)
    const StepCursor sc = *step_src.cur;
    coef = step_src.coefs[sc.ctr];
    seed = sc.seed;
    stream_id = sc.stream_base + (uint64_t)sc.ctr;
    if (z) z += (int64_t)(sc.ctr - sc.noise_step0) * (M * C);
  }
  float* hs = sm;                    // [UT][d+1]
  float* ws = sm + UT * (d + 1);     // [C][d]
  const int64_t row0 = (int64_t)blockIdx.x * UT;
  const int rows = (int)((M - row0 < UT) ? (M - row0) : UT);
  const int d4 = d / 4;
  const float4* src = reinterpret_cast<const float4*>(h + row0 * d);
  // six 16-byte loads per thread in flight before the first shared-memory store (the kernel is a pure HBM read)
  const int n4 = rows * d4;
  for (int i0 = threadIdx.x; i0 < n4; i0 += 6 * UT) {
    float4 v[6];
#pragma unroll
    for (int u = 0; u < 6; ++u)
      if (i0 + u * UT < n4) v[u] = __ldcs(src + i0 + u * UT);
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int i = i0 + u * UT;
      if (i < n4) {
        const int r = i / d4, c = (i - r * d4) * 4;
        float* dst = hs + r * (d + 1) + c;
        dst[0] = v[u].x, dst[1] = v[u].y, dst[2] = v[u].z, dst[3] = v[u].w;
      }
    }
  }
  for (int i = threadIdx.x; i < C * d; i += UT) ws[i] = Wu[i];
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= rows) return;
  const int64_t row = row0 + r;
  const float g = do_step ? G[row % L] : 0.f;
  const float* hr = hs + r * (d + 1);
  for (int c = 0; c < C; ++c) {
    float p0 = 0.f, p1 = 0.f;
    const float* w = ws + c * d;
    int k = 0;
    for (; k + 1 < d; k += 2) {
      p0 = fmaf(hr[k], w[k], p0);
      p1 = fmaf(hr[k + 1], w[k + 1], p1);
    }
    if (k < d) p0 = fmaf(hr[k], w[k], p0);
    const float p = (p0 + p1) + bu[c];
    const int64_t e = row * C + c;
    if (score_out) score_out[e] = p;
    if (do_step) {
      float zz;
      if (z) {
        zz = z[e];
      } else {
        const float4 n4 = philox_normal4(seed, stream_id, (uint64_t)e >> 2);
        const int q = (int)(e & 3);
        zz = q == 0 ? n4.x : q == 1 ? n4.y : q == 2 ? n4.z : n4.w;
      }
      x_out[e] = sde_update(coef, g, x[e], p, zz);
    }
  }
}

int launch_unembed_step(const float* h, const float* Wu, const float* bu, float* score_out, int do_step,
                        const ffd_step_coef* coef, const float* G, const float* x, const float* z, uint64_t seed,
                        uint64_t stream_id, float* x_out, int64_t M, int L, int C, int d, cudaStream_t s,
                        StepSource src) {
  if (d > 256) {
    set_error("d_model %d > 256 unsupported", d);
    return FFD_ERR_UNSUPPORTED;
  }
  ffd_step_coef c = {};
  if (coef) c = *coef;
  const size_t smem = sizeof(float) * ((size_t)UT * (d + 1) + (size_t)C * d);
  if (d % 4 == 0 && (uintptr_t)h % 16 == 0 && smem <= 96 * 1024) {
    if (smem > 48 * 1024)
      FFD_CHECK_CUDA(cudaFuncSetAttribute(unembed_step_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    unembed_step_tiled_kernel<<<(unsigned)((M + UT - 1) / UT), UT, smem, s>>>(h, Wu, bu, score_out, do_step, c, G, x,
                                                                              z, seed, stream_id, x_out, M, L, C, d, src);
    FFD_CHECK_LAUNCH("unembed_step_tiled_kernel");
    return FFD_OK;
  }
  unembed_step_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(h, Wu, bu, score_out, do_step, c, G, x, z, seed,
                                                              stream_id, x_out, M, L, C, d, src);
  FFD_CHECK_LAUNCH("unembed_step_kernel");
  return FFD_OK;
}

__global__ void step_advance_kernel(StepCursor* cur) { cur->ctr += 1; }

int launch_step_advance(StepCursor* cur, cudaStream_t s) {
  step_advance_kernel<<<1, 1, 0, s>>>(cur);
  FFD_CHECK_LAUNCH("step_advance_kernel");
  return FFD_OK;
}

// stand-alone SDE update (VPScheduler.step / VEScheduler.step called on their own)
__global__ void sde_step_kernel(ffd_step_coef coef, const float* __restrict__ G, const float* __restrict__ x,
                                const float* __restrict__ score, const float* __restrict__ z, uint64_t seed,
                                uint64_t stream_id, float* __restrict__ x_out, int64_t n, int L, int C) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // element quad
  if (q * 4 >= n) return;
  float zz[4];
  if (!z) {
    const float4 n4 = philox_normal4(seed, stream_id, (uint64_t)q);
    zz[0] = n4.x; zz[1] = n4.y; zz[2] = n4.z; zz[3] = n4.w;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t e = q * 4 + r;
    if (e >= n) break;
    const int l = (int)((e / C) % L);
    x_out[e] = sde_update(coef, G[l], x[e], score[e], z ? z[e] : zz[r]);
  }
}

int launch_sde_step(const ffd_step_coef* c, const float* G, const float* x, const float* score, const float* z,
                    uint64_t seed, uint64_t stream_id, float* x_out, int64_t B, int L, int C, cudaStream_t s) {
  const int64_t n = B * L * C, nq = (n + 3) / 4;
  if (n == 0) return FFD_OK;
  sde_step_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(*c, G, x, score, z, seed, stream_id, x_out, n, L,
                                                               C);
  FFD_CHECK_LAUNCH("sde_step_kernel");
  return FFD_OK;
}

// prior: x = scale * (G[l] * z)   (sde.py:79-87 G_matrix @ z; VE :125-127 scale = sigma_max)
__global__ void prior_kernel(const float* __restrict__ G, const float* __restrict__ z, uint64_t seed,
                             uint64_t stream_id, float scale, int use_scale, float* __restrict__ x_out,
                             int64_t n, int L, int C) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  float zz[4];
  if (!z) {
    const float4 n4 = philox_normal4(seed, stream_id, (uint64_t)q);
    zz[0] = n4.x; zz[1] = n4.y; zz[2] = n4.z; zz[3] = n4.w;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t e = q * 4 + r;
    if (e >= n) break;
    const int l = (int)((e / C) % L);
    float v = __fmul_rn(G[l], z ? z[e] : zz[r]);
    if (use_scale) v = __fmul_rn(scale, v);
    x_out[e] = v;
  }
}

int launch_prior(const float* G, const float* z, uint64_t seed, uint64_t stream_id, float scale, float* x_out,
                 int64_t B, int L, int C, cudaStream_t s) {
  const int64_t n = B * L * C, nq = (n + 3) / 4;
  if (n == 0) return FFD_OK;
  prior_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(G, z, seed, stream_id, scale, scale != 1.0f, x_out, n,
                                                            L, C);
  FFD_CHECK_LAUNCH("prior_kernel");
  return FFD_OK;
}

__global__ void philox_normal_kernel(uint64_t seed, uint64_t stream_id, float* __restrict__ out, int64_t n) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  const float4 n4 = philox_normal4(seed, stream_id, (uint64_t)q);
  const float v[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (q * 4 + r < n) out[q * 4 + r] = v[r];
}

int launch_philox_normal(uint64_t seed, uint64_t stream_id, float* out, int64_t n, cudaStream_t s) {
  const int64_t nq = (n + 3) / 4;
  if (n == 0) return FFD_OK;
  philox_normal_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(seed, stream_id, out, n);
  FFD_CHECK_LAUNCH("philox_normal_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// embeddings
// ------------------------------------------------------------------------------------
// nn.Embedding(max_norm): rows with ||row||_2 > max_norm are scaled by max_norm/(norm+1e-7)
// (ATen embedding_renorm_; transformer.py:13-15,27).  One warp per row.
__global__ void posemb_renorm_kernel(float* __restrict__ table, int L, int d, float max_norm) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= L) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = table[(int64_t)row * d + k];
    s = fmaf(v, v, s);
  }
  const float norm = sqrtf(warp_sum(s));
  if (norm > max_norm) {
    const float sc = max_norm / (norm + 1e-7f);
    for (int k = lane; k < d; k += 32) table[(int64_t)row * d + k] *= sc;
  }
}

int launch_posemb_renorm(float* table, int L, int d, float max_norm, cudaStream_t s) {
  posemb_renorm_kernel<<<(L + 7) / 8, 256, 0, s>>>(table, L, d, max_norm);
  FFD_CHECK_LAUNCH("posemb_renorm_kernel");
  return FFD_OK;
}

// GaussianFourierProjection (transformer.py:80-91): proj = ((t*W)*2)*pi in fp32,
// e = [sin(proj), cos(proj)][:d], out = dense_w @ e + dense_b.  One CTA per timestep.
__global__ void time_embedding_kernel(const float* __restrict__ t, const float* __restrict__ W,
                                      const float* __restrict__ dense_w, const float* __restrict__ dense_b, int d,
                                      float* __restrict__ out) {
  extern __shared__ float e[];  // [d]
  const int i = blockIdx.x;
  const int half = (d + 1) / 2;
  const float ti = t[i];
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const int w = k < half ? k : k - half;
    const float proj = __fmul_rn(__fmul_rn(__fmul_rn(ti, W[w]), 2.0f), 3.14159265358979323846f);
    e[k] = k < half ? sinf(proj) : cosf(proj);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(dense_w[j * d + k], e[k], acc);
    out[(int64_t)i * d + j] = acc + dense_b[j];
  }
}

int launch_time_embedding(const float* t, int n, const float* W, const float* dense_w, const float* dense_b, int d,
                          float* out, cudaStream_t s) {
  if (n == 0) return FFD_OK;
  time_embedding_kernel<<<n, 128, d * sizeof(float), s>>>(t, W, dense_w, dense_b, d, out);
  FFD_CHECK_LAUNCH("time_embedding_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// boundary DFT as a dense real matrix (fourier.py:8-94; matrix form SURVEY a17)
// ------------------------------------------------------------------------------------
__global__ void dft_matrix_kernel(float* __restrict__ out, int L, int inverse) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * L) return;
  const int nr = (L + 2) / 2;  // ceil((L+1)/2)
  int row = idx / L, col = idx % L;
  // U[k, n]; the inverse is (diag(w) U)^T so out[n, k] = w[k] U[k, n]
  const int k = inverse ? col : row, n = inverse ? row : col;
  const int freq = k < nr ? k : k - (nr - 1);
  const int ph = (int)(((int64_t)freq * n) % L);  // exact phase reduction
  const double ang = 2.0 * (double)ph / (double)L;
  double v = (k < nr ? cospi(ang) : -sinpi(ang)) / sqrt((double)L);
  if (inverse) {
    const bool single = (k == 0) || (L % 2 == 0 && k == L / 2);
    if (!single) v *= 2.0;
  }
  out[idx] = (float)v;
}

int launch_dft_matrix(float* out, int L, int inverse, cudaStream_t s) {
  dft_matrix_kernel<<<(L * L + 255) / 256, 256, 0, s>>>(out, L, inverse);
  FFD_CHECK_LAUNCH("dft_matrix_kernel");
  return FFD_OK;
}

// y[b,k,c] = sum_n T[k,n] * (x[b,n,c]*scale[n,c] + shift[n,c]); one CTA per sample
__global__ void __launch_bounds__(256)
dft_apply_kernel(const float* __restrict__ T, const float* __restrict__ x, const float* __restrict__ scale,
                 const float* __restrict__ shift, float* __restrict__ y, int L, int C) {
  extern __shared__ float xs[];  // [L*C]
  const int64_t b = blockIdx.x;
  for (int e = threadIdx.x; e < L * C; e += blockDim.x) {
    float v = x[b * L * C + e];
    if (scale) v = __fmul_rn(v, scale[e]);
    if (shift) v = __fadd_rn(v, shift[e]);
    xs[e] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * C; e += blockDim.x) {
    const int k = e / C, c = e % C;
    const float* Tk = T + (int64_t)k * L;
    float acc = 0.f;
    for (int n = 0; n < L; ++n) acc = fmaf(Tk[n], xs[n * C + c], acc);
    y[b * L * C + e] = acc;
  }
}

// The same product for a GROUP of samples per CTA, so that every element of T is read from shared memory once per 16
// output columns instead of once per column from L2 (the one-sample kernel streams the whole L x L matrix per sample:
// 2.3 GB of L2 reads for an ECG batch of 16384).  A CTA owns NB = 16 / C samples = up to 16 columns j = (b - b0) C + c;
// thread k accumulates y[., k, .] for all of them over n ASCENDING with fmaf -- the same rounding order as
// dft_apply_kernel, so both kernels return identical bits.  T enters in [256 k x 32 n] tiles (coalesced rows).
constexpr int DFT_COLS = 16;
constexpr int DFT_KT = 256;
__global__ void __launch_bounds__(DFT_KT)
dft_apply_tiled_kernel(const float* __restrict__ T, const float* __restrict__ x, const float* __restrict__ scale,
                       const float* __restrict__ shift, float* __restrict__ y, int64_t B, int L, int C, int NB) {
  extern __shared__ float sm[];
  float* xs = sm;                                  // [L][DFT_COLS]
  float* Ts = sm + (size_t)L * DFT_COLS;           // [DFT_KT][33]
  const int64_t b0 = (int64_t)blockIdx.x * NB;
  const int nb = (int)(B - b0 < NB ? B - b0 : NB);
  const int cols = nb * C;
  for (int e = threadIdx.x; e < L * DFT_COLS; e += DFT_KT) {
    const int n = e / DFT_COLS, j = e % DFT_COLS;
    float v = 0.f;
    if (j < cols) {
      const int bb = j / C, c = j % C;
      v = x[((b0 + bb) * L + n) * C + c];
      if (scale) v = __fmul_rn(v, scale[n * C + c]);
      if (shift) v = __fadd_rn(v, shift[n * C + c]);
    }
    xs[e] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k0 = 0; k0 < L; k0 += DFT_KT) {
    const int k = k0 + threadIdx.x;
    float acc[DFT_COLS];
#pragma unroll
    for (int j = 0; j < DFT_COLS; ++j) acc[j] = 0.f;
    for (int n0 = 0; n0 < L; n0 += 32) {
      __syncthreads();  // xs written / previous tile consumed
      for (int r = warp; r < DFT_KT; r += DFT_KT / 32) {
        const int kk = k0 + r, n = n0 + lane;
        Ts[r * 33 + lane] = (kk < L && n < L) ? T[(int64_t)kk * L + n] : 0.f;
      }
      __syncthreads();
      const int nn_end = L - n0 < 32 ? L - n0 : 32;
      for (int nn = 0; nn < nn_end; ++nn) {
        const float t = Ts[threadIdx.x * 33 + nn];
        const float4* xr = reinterpret_cast<const float4*>(xs + (size_t)(n0 + nn) * DFT_COLS);
#pragma unroll
        for (int q = 0; q < DFT_COLS / 4; ++q) {
          const float4 v = xr[q];
          acc[4 * q] = fmaf(t, v.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(t, v.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(t, v.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(t, v.w, acc[4 * q + 3]);
        }
      }
    }
    if (k < L) {
#pragma unroll
      for (int j = 0; j < DFT_COLS; ++j)
        if (j < cols) y[((b0 + j / C) * L + k) * C + j % C] = acc[j];
    }
  }
}

int launch_dft_apply(const float* T, const float* x, const float* scale, const float* shift, float* y, int64_t B,
                     int L, int C, cudaStream_t s) {
  if (B == 0) return FFD_OK;
  if (C <= DFT_COLS && B >= 64) {
    const int NB = DFT_COLS / C;
    const size_t smem = ((size_t)L * DFT_COLS + (size_t)DFT_KT * 33) * sizeof(float);
    if (smem <= 200 * 1024) {
      if (smem > 48 * 1024)
        FFD_CHECK_CUDA(cudaFuncSetAttribute(dft_apply_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
      dft_apply_tiled_kernel<<<(unsigned)((B + NB - 1) / NB), DFT_KT, smem, s>>>(T, x, scale, shift, y, B, L, C, NB);
      FFD_CHECK_LAUNCH("dft_apply_tiled_kernel");
      return FFD_OK;
    }
  }
  const size_t smem = (size_t)L * C * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("L*C = %d too large for the boundary DFT kernel", L * C);
    return FFD_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    FFD_CHECK_CUDA(cudaFuncSetAttribute(dft_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dft_apply_kernel<<<(unsigned)B, 256, smem, s>>>(T, x, scale, shift, y, L, C);
  FFD_CHECK_LAUNCH("dft_apply_kernel");
  return FFD_OK;
}

// ------------------------------------------------------------------------------------
// FreSca score filter (src/fdiff/utils/fresca.py:111-268): rfft along L, low band x l, high band x h, irfft.
// In the real-packed DFT basis (fourier.py layout) that is a per-frequency gain between two dense transforms.
// ------------------------------------------------------------------------------------
// part[block, k] = sum over this block's samples and all channels of |X_k| (packed layout: Re at row k, Im at
// nr-1+k).  No atomics: the partial sums are reduced in a fixed order by fresca_gain_kernel, so that the band
// boundary the "energy" strategy derives from them is the same in every run.
constexpr int FRESCA_BLOCKS = 592;
__global__ void __launch_bounds__(256)
fresca_spectrum_kernel(const float* __restrict__ xt, float* __restrict__ part, int64_t B, int L, int C, int n_freq) {
  const int nr = (L + 2) / 2;  // = n_freq
  const int64_t per = (B + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = (int64_t)blockIdx.x * per, b1 = b0 + per < B ? b0 + per : B;
  for (int k = threadIdx.x; k < n_freq; k += blockDim.x) {
    const bool has_im = k >= 1 && (nr - 1 + k) < L;
    float acc = 0.f;
    for (int64_t b = b0; b < b1; ++b)
      for (int c = 0; c < C; ++c) {
        const float re = xt[(b * L + k) * C + c];
        const float im = has_im ? xt[(b * L + nr - 1 + k) * C + c] : 0.f;
        acc += sqrtf(re * re + im * im);
      }
    part[(int64_t)blockIdx.x * n_freq + k] = acc;
  }
}

// cutoff index and per-row gains.  strategy 0 = spatial (k <= r0 * n_freq), 1 = energy: first i whose cumulative
// mean magnitude reaches r0 * total (fresca.py:41-60); gain[row, c] = l for rows of frequencies <= Rc, else h.
// n_part > 0: spec[k] is first reduced from the per-block partial spectra, blocks in ascending order.
__global__ void fresca_gain_kernel(float* __restrict__ spec, const float* __restrict__ part, int n_part,
                                   float* __restrict__ gain, int L, int C, int n_freq, float low, float high,
                                   float cutoff_ratio, int strategy, float inv_count) {
  __shared__ int rc_s;
  if (strategy == 1) {
    for (int k = threadIdx.x; k < n_freq; k += blockDim.x) {
      float acc = 0.f;
      for (int b = 0; b < n_part; ++b) acc += part[(int64_t)b * n_freq + k];
      spec[k] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int rc = 0;
    if (strategy == 0) {
      const float lim = cutoff_ratio * (float)n_freq;
      rc = -1;
      for (int k = 0; k < n_freq; ++k)
        if ((float)k <= lim) rc = k;
    } else {
      float etot = 0.f;  // torch sums the fp32 spectrum in fp32
      for (int k = 0; k < n_freq; ++k) etot += spec[k] * inv_count;
      double cum = 0.0;  // the reference accumulates Python floats
      const double thr = (double)cutoff_ratio * (double)etot;
      for (int k = 0; k < n_freq; ++k) {
        cum += (double)(spec[k] * inv_count);
        if (cum >= thr) {
          rc = k;
          break;
        }
      }
    }
    rc_s = rc;
  }
  __syncthreads();
  const int rc = rc_s, nr = n_freq;
  for (int i = threadIdx.x; i < L * C; i += blockDim.x) {
    const int row = i / C;
    const int k = row < nr ? row : row - (nr - 1);
    gain[i] = k <= rc ? low : high;
  }
}

int64_t fresca_scratch_floats(int64_t B, int L, int C) {
  return B * L * C + (int64_t)L * C + (int64_t)(FRESCA_BLOCKS + 1) * (L / 2 + 1);
}

int launch_fresca(const float* U, const float* V, const float* score, float* out, float* scratch, int64_t B, int L,
                  int C, float low, float high, float cutoff_ratio, int strategy, cudaStream_t s) {
  // scratch: [B*L*C] transformed scores, [L*C] gains, [n_freq] spectrum, [FRESCA_BLOCKS][n_freq] partial spectra
  const int n_freq = L / 2 + 1;
  float* xt = scratch;
  float* gain = scratch + B * L * C;
  float* spec = gain + (int64_t)L * C;
  float* part = spec + n_freq;
  int rc;
  if ((rc = launch_dft_apply(U, score, nullptr, nullptr, xt, B, L, C, s))) return rc;
  unsigned blocks = 0;
  if (strategy == 1) {
    blocks = (unsigned)(B < FRESCA_BLOCKS ? B : FRESCA_BLOCKS);
    fresca_spectrum_kernel<<<blocks, 256, 0, s>>>(xt, part, B, L, C, n_freq);
    FFD_CHECK_LAUNCH("fresca_spectrum_kernel");
  }
  fresca_gain_kernel<<<1, 256, 0, s>>>(spec, part, (int)blocks, gain, L, C, n_freq, low, high, cutoff_ratio, strategy,
                                       1.0f / (float)(B * C));
  FFD_CHECK_LAUNCH("fresca_gain_kernel");
  return launch_dft_apply(V, xt, gain, nullptr, out, B, L, C, s);
}

// ---- E2-CRF event intensity (caching.py:541-557) ---------------------------------------------------
// One block; warp w owns rows w, w + 32, ...: energy[row] = || crf[row] - prev[row] ||_2, then the mean over rows in a
// fixed order (deterministic), intensity = min(1, mean / tau_0).
__global__ void event_intensity_kernel(const float* __restrict__ crf, const float* __restrict__ prev, int64_t rows, int d,
                                       float tau_0, float* __restrict__ out) {
  __shared__ float part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int64_t r = warp; r < rows; r += 32) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float dl = fabsf(crf[r * d + c] - prev[r * d + c]);
      ss = fmaf(dl, dl, ss);
    }
    ss = warp_sum(ss);
    acc += sqrtf(ss);
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float v = part[lane];
    v = warp_sum(v);
    if (lane == 0) {
      const float avg = v / (float)rows;
      out[0] = avg;
      out[1] = fminf(1.f, avg / tau_0);
    }
  }
}

int launch_event_intensity(const float* crf, const float* prev, int64_t rows, int d, float tau_0, float* out,
                           cudaStream_t s) {
  event_intensity_kernel<<<1, 1024, 0, s>>>(crf, prev, rows, d, tau_0, out);
  FFD_CHECK_LAUNCH("event_intensity_kernel");
  return FFD_OK;
}

}  // namespace ffd
