// E2-CRF event-driven schedule decided on the device (include/ffd_b200.h: ffd_event_args).
//
// The schedule of the reference's paper (submission/TSD_Cache_Paper_Final.pdf section 3.3: Eq. 14 event intensity,
// Eq. 15 recompute set with energy-weighted thresholds, Eq. 17-18 error-feedback correction, Algorithm 1) -- the
// code the reference ships decides by step index only (caching.py:131-181).  Two kernels, both HBM/latency-sized
// work on sample 0's rows only ([L, d] fp32, 54 KB for ECG), so one block each:
//   event_decide_kernel      one block: per-token change of the final-layer CRF, intensity r, accumulators, the
//                            blend weights of the next step's refresh, counters.  Fixed reduction order.
//   event_refresh_kv_kernel  one block per token, gated by the weight on the device: K/V rows of one layer of the
//                            sample-0 cache from that layer's input (the mixed branch's projection,
//                            cached_transformer.py:282-296), blended into the cache.
#include "ffd_internal.h"

namespace ffd {

// uniform in (0, 1) for token k of step `step`: Philox-4x32-10, counter (k, 0, step_lo, step_hi), key = seed
__device__ __forceinline__ float event_uniform(uint64_t seed, uint32_t step, uint32_t k) {
  const uint4 r = philox4x32_10(make_uint4(k, 0u, step, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return ((float)r.x + 0.5f) * 2.3283064365386963e-10f;
}

__global__ void __launch_bounds__(1024, 1)
event_decide_kernel(ffd_event_args ev, const float* __restrict__ z, const float* __restrict__ x0, int L, int d, int C,
                    int step, int first) {
  extern __shared__ float sh[];  // delta[L], num[32], den[32]
  float* delta = sh;
  float* pnum = sh + L;
  float* pden = pnum + 32;
  __shared__ int cnt[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  if (first) {
    for (int i = threadIdx.x; i < L * d; i += blockDim.x) ev.z_prev[i] = z[i];
    for (int k = threadIdx.x; k < L; k += blockDim.x) ev.accum[k] = 0.f, ev.weight[k] = 0.f;
    if (threadIdx.x < FFD_EVENT_STATS) ev.stats[threadIdx.x] = 0.f;
    return;
  }
  // pass 1: warp w owns tokens w, w + 32, ...
  float num = 0.f, den = 0.f;
  for (int k = warp; k < L; k += 32) {
    float ss = 0.f, nn = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float p = ev.z_prev[k * d + c];
      const float dl = z[k * d + c] - p;
      ss = fmaf(dl, dl, ss);
      nn = fmaf(p, p, nn);
    }
    ss = warp_sum(ss);
    nn = warp_sum(nn);
    if (lane == 0) delta[k] = sqrtf(ss);
    num += ss;
    den += nn;
  }
  if (lane == 0) pnum[warp] = num, pden[warp] = den;
  __syncthreads();
  // every thread forms r in the same order
  float tn = 0.f, td = 0.f;
  for (int w = 0; w < 32; ++w) tn += pnum[w], td += pden[w];
  const float r = tn / (td + ev.eta);                                    // Eq. 14
  const bool feedback = ((step + 1) % ev.R == 0) || r > ev.tau_warn;     // Algorithm 1, line 16
  const float gamma = fminf(0.1f, 0.5f * r);                             // section 4: gamma = min(0.1, r / 2)
  // pass 2: the next step's weights
  int n_rec = 0, n_cor = 0;
  for (int k = threadIdx.x; k < L; k += blockDim.x) {
    float e = 0.f;
    for (int c = 0; c < C; ++c) e = fmaf(x0[k * C + c], x0[k * C + c], e);
    const float tau_k = ev.tau_0 / (ev.eps + e);                          // Eq. 15: tau_0 (eps + ||x_k||^2)^-1
    const float a = (ev.accumulate ? ev.accum[k] : 0.f) + delta[k];
    const bool probe = k > ev.K && ev.probe_ratio > 0.f && event_uniform(ev.seed, (uint32_t)step, (uint32_t)k) < ev.probe_ratio;
    const bool rec = k <= ev.K || a > tau_k || probe;
    const float w = rec ? 1.f : (feedback ? gamma : 0.f);
    ev.weight[k] = w;
    ev.accum[k] = rec ? 0.f : a;
    n_rec += rec ? 1 : 0;
    n_cor += (!rec && w > 0.f) ? 1 : 0;
  }
  if (n_rec) atomicAdd(&cnt[0], n_rec);
  if (n_cor) atomicAdd(&cnt[1], n_cor);
  for (int i = threadIdx.x; i < L * d; i += blockDim.x) ev.z_prev[i] = z[i];  // pass 1 is behind the barrier above
  __syncthreads();
  if (threadIdx.x == 0) {
    ev.stats[0] += 1.f;
    ev.stats[1] += (float)cnt[0];
    ev.stats[2] += (float)cnt[1];
    ev.stats[3] += feedback ? 1.f : 0.f;
    ev.stats[4] = r;
    ev.stats[5] += r;
  }
}

int launch_event_decide(const ffd_event_args& ev, const float* z, const float* x0, int L, int d, int C, int step,
                        int first, cudaStream_t s) {
  event_decide_kernel<<<1, 1024, (size_t)(L + 64) * sizeof(float), s>>>(ev, z, x0, L, d, C, step, first);
  FFD_CHECK_LAUNCH("event_decide_kernel");
  return FFD_OK;
}

// Wkv = rows d .. 3d of in_proj_weight ([2d, d]), bkv likewise: thread j < 2d owns output feature j of token k.
__global__ void event_refresh_kv_kernel(const float* __restrict__ Wkv, const float* __restrict__ bkv,
                                        const float* __restrict__ h0, const float* __restrict__ weight,
                                        float* __restrict__ kc, float* __restrict__ vc, int L, int d, int hd) {
  extern __shared__ float row[];
  const int k = blockIdx.x;
  const float w = weight[k];
  if (!(w > 0.f)) return;  // reuse: decided on the device, the host never learns which
  for (int c = threadIdx.x; c < d; c += blockDim.x) row[c] = h0[(int64_t)k * d + c];
  __syncthreads();
  const int j = threadIdx.x;
  if (j >= 2 * d) return;
  float acc = bkv[j];
  const float* wr = Wkv + (int64_t)j * d;
  for (int c = 0; c < d; ++c) acc = fmaf(wr[c], row[c], acc);
  float* dst = j < d ? kc : vc;
  const int jj = j < d ? j : j - d;
  const int64_t idx = ((int64_t)(jj / hd) * L + k) * hd + jj % hd;  // [H, L, hd] slice of this layer
  dst[idx] = w >= 1.f ? acc : fmaf(w, acc - dst[idx], dst[idx]);   // Eq. 18: z^ <- z^ + gamma (z - z^)
}

int launch_event_refresh_kv(const float* Wkv, const float* bkv, const float* h0, const float* weight, float* kc,
                            float* vc, int L, int d, int H, cudaStream_t s) {
  const int threads = (2 * d + 31) / 32 * 32;
  event_refresh_kv_kernel<<<L, threads, (size_t)d * sizeof(float), s>>>(Wkv, bkv, h0, weight, kc, vc, L, d, d / H);
  FFD_CHECK_LAUNCH("event_refresh_kv_kernel");
  return FFD_OK;
}

}  // namespace ffd
