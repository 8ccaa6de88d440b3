// Sample-quality metrics on the device: sliced / marginal Wasserstein-2 distances between two
// sample sets, fdiff/sampling/metrics.py:101-216 over fdiff/utils/wasserstein.py:96-199.  The reference projects
// with numpy and calls POT's ot.emd2_1d per direction on the host; here the [n_dir, N] projections, the per-direction
// sort and the exact 1-D optimal-transport cost all stay on the GPU, so a 10^4..10^5-sample run is scored without
// copying the samples back.  HBM / shared-memory bound integer-and-compare work: no tensor cores.
#include "ffd_internal.h"

namespace ffd {

// out[r, n] = sum_k data[n, k] * dirs[r, k]      (wasserstein.py:146-149 `_project`, all directions at once)
// 64 x 64 output tile per block, 16 x 16 threads, 4 x 4 outputs per thread, K staged through shared memory.
__global__ void metric_project_kernel(const float* __restrict__ data, int64_t N, int D, const float* __restrict__ dirs,
                                      int R, float* __restrict__ out) {
  __shared__ float sa[16][65], sb[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t n0 = (int64_t)blockIdx.x * 64;
  const int r0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < D; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int row = i >> 4, k = i & 15;
      sa[k][row] = (n0 + row < N && k0 + k < D) ? data[(n0 + row) * D + k0 + k] : 0.f;
      sb[k][row] = (r0 + row < R && k0 + k < D) ? dirs[(int64_t)(r0 + row) * D + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[k][tx + 16 * i], b[i] = sb[k][ty + 16 * i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(a[i], b[j], acc[j][i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 16 * j;
      const int64_t n = n0 + tx + 16 * i;
      if (r < R && n < N) out[(int64_t)r * N + n] = acc[j][i];
    }
}

// ascending sort of every row of x[R, N] in place: one block per row, bitonic network in shared memory
// (N <= 32768: the row, padded with +inf to a power of two, fits in 128 KB).
__global__ void metric_sort_rows_kernel(float* __restrict__ x, int64_t N, int P) {
  extern __shared__ float sv[];
  float* row = x + (int64_t)blockIdx.x * N;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sv[i] = i < N ? row[i] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const float a = sv[i], b = sv[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            sv[i] = b;
            sv[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) row[i] = sv[i];
}

// deterministic block sum of doubles (fixed tree)
__device__ double block_sum(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// out[r] = sqrt(emd2_1d(a[r, :], b[r, :])) for ascending rows with uniform weights 1/N and 1/M and squared-Euclidean
// ground cost (wasserstein.py:117-119,141-143; POT's emd_1d on sorted inputs): the optimal 1-D plan couples equal
// quantiles.  In units of 1/(N M), a_i holds the mass interval [i M, (i + 1) M) and b_j holds [j N, (j + 1) N):
// thread i adds (a_i - b_j)^2 x overlap for the few j its interval meets.  One block per row, fp64 accumulation.
__global__ void metric_w2_sorted_kernel(const float* __restrict__ a, int64_t N, const float* __restrict__ b, int64_t M,
                                        float* __restrict__ out) {
  __shared__ double sh[256];
  const float* ar = a + (int64_t)blockIdx.x * N;
  const float* br = b + (int64_t)blockIdx.x * M;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const int64_t lo = i * M, hi = (i + 1) * M;
    const double ai = (double)ar[i];
    for (int64_t j = lo / N; j * N < hi; ++j) {
      const int64_t s = j * N > lo ? j * N : lo, e = (j + 1) * N < hi ? (j + 1) * N : hi;
      const double df = ai - (double)br[j];
      acc += df * df * (double)(e - s);
    }
  }
  const double tot = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)sqrt(tot / ((double)N * (double)M));
}

// population standard deviation of every row (np.std, wasserstein.py:157-159 "standardise")
__global__ void metric_row_std_kernel(const float* __restrict__ x, int64_t N, float* __restrict__ out) {
  __shared__ double sh[256];
  const float* r = x + (int64_t)blockIdx.x * N;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s += (double)r[i];
  const double mean = block_sum(s, sh) / (double)N;
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double dlt = (double)r[i] - mean;
    q += dlt * dlt;
  }
  const double var = block_sum(q, sh) / (double)N;
  if (threadIdx.x == 0) out[blockIdx.x] = (float)sqrt(var);
}

// out[b, k, c] = Re(X_k)^2 + Im(X_k)^2 from the real-packed spectrum xt[b, :, c] (fourier.py:108-131): real parts in
// rows 0..n_r-1, imaginary parts of bins 1.. in rows n_r.. (bin 0, and the Nyquist bin of an even L, have none).
__global__ void spectral_density_kernel(const float* __restrict__ xt, float* __restrict__ out, int64_t B, int L, int C) {
  const int nr = L / 2 + 1;
  const int64_t total = B * nr * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int k = (int)((i / C) % nr);
    const int64_t b = i / ((int64_t)C * nr);
    const float re = xt[(b * L + k) * C + c];
    const bool has_im = k > 0 && !(L % 2 == 0 && k == nr - 1);
    const float im = has_im ? xt[(b * L + nr + k - 1) * C + c] : 0.f;
    out[i] = re * re + im * im;
  }
}

int launch_spectral_density(const float* xt, float* out, int64_t B, int L, int C, cudaStream_t s) {
  const int64_t total = B * (L / 2 + 1) * C;
  const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  spectral_density_kernel<<<grid ? grid : 1, 256, 0, s>>>(xt, out, B, L, C);
  FFD_CHECK_LAUNCH("spectral_density_kernel");
  return FFD_OK;
}

int launch_metric_project(const float* data, int64_t N, int D, const float* dirs, int R, float* out, cudaStream_t s) {
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((R + 63) / 64));
  metric_project_kernel<<<grid, 256, 0, s>>>(data, N, D, dirs, R, out);
  FFD_CHECK_LAUNCH("metric_project_kernel");
  return FFD_OK;
}

int launch_metric_sort_rows(float* x, int R, int64_t N, cudaStream_t s) {
  int P = 1;
  while (P < N) P <<= 1;
  const int smem = P * 4;
  FFD_CHECK_CUDA(cudaFuncSetAttribute(metric_sort_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  metric_sort_rows_kernel<<<R, 1024, smem, s>>>(x, N, P);
  FFD_CHECK_LAUNCH("metric_sort_rows_kernel");
  return FFD_OK;
}

int launch_metric_w2_sorted(const float* a, int64_t N, const float* b, int64_t M, int R, float* out, cudaStream_t s) {
  metric_w2_sorted_kernel<<<R, 256, 0, s>>>(a, N, b, M, out);
  FFD_CHECK_LAUNCH("metric_w2_sorted_kernel");
  return FFD_OK;
}

int launch_metric_row_std(const float* x, int R, int64_t N, float* out, cudaStream_t s) {
  metric_row_std_kernel<<<R, 256, 0, s>>>(x, N, out);
  FFD_CHECK_LAUNCH("metric_row_std_kernel");
  return FFD_OK;
}

}  // namespace ffd
