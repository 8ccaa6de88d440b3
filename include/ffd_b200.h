/*
 * ffd_b200.h -- C ABI of the B200-native reverse-SDE sampling path.
 *
 * The reference (NoakLiu/FastFourierDiffusion) has no FFI: its hot path is a
 * Python class API over stock ATen ops.  Each entry point below replaces the
 * ATen op sequence of one reference function and is what a maintainer would
 * bind (ctypes) underneath that function:
 *
 *   ffd_dft_matrix / ffd_dft_apply   src/fdiff/utils/fourier.py:8-52 (dft), :55-94 (idft)
 *   ffd_fresca                       src/fdiff/utils/fresca.py:111-268 (apply_fresca_to_score)
 *   ffd_posemb_renorm                src/fdiff/models/transformer.py:13-15,26-28 (Embedding max_norm)
 *   ffd_time_embedding               src/fdiff/models/transformer.py:77-91 (GaussianFourierProjection)
 *   ffd_score_forward                src/fdiff/models/score_models.py:79-194 (ScoreModule.forward,
 *                                    _forward_with_cache) and
 *                                    src/fdiff/models/cached_transformer.py:106-329 (all five branches),
 *                                    src/fdiff/utils/caching.py:251-396 (KV store), :459-480 (CRF store)
 *   ffd_sde_step                     src/fdiff/schedulers/sde.py:129-165 (VE), :215-246 (VP)
 *   ffd_prior                        src/fdiff/schedulers/sde.py:79-87, :125-127
 *   ffd_philox_normal                torch.randn / randn_like call sites sde.py:85,157,238
 *   ffd_sample_steps                 src/fdiff/sampling/sampler.py:157-210 (the step loop)
 *   ffd_event_decide / _refresh_kv   the paper's Algorithm 1 (no counterpart in the reference's code)
 *
 * Conventions: plain pointers and sizes only; all pointers are DEVICE pointers
 * unless the name ends in _host; the library never allocates, never frees and
 * never synchronises -- every call only enqueues work on `stream`.  Return
 * value 0 = ok, negative = error (ffd_last_error() gives the message for the
 * calling thread).  Tensors are contiguous row-major fp32 in the reference's
 * own layouts: X[B, L, C], weights as in ScoreModule.state_dict().
 */
#ifndef FFD_B200_H
#define FFD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFD_OK 0
#define FFD_ERR_ARG (-1)       /* bad argument (shape, null pointer, unsupported size) */
#define FFD_ERR_CUDA (-2)      /* a CUDA call or launch failed */
#define FFD_ERR_WORKSPACE (-3) /* workspace too small */
#define FFD_ERR_UNSUPPORTED (-4)

/* arithmetic of the score-model GEMMs */
#define FFD_PREC_FP32 0 /* CUDA-core fp32 FMA everywhere */
#define FFD_PREC_FP16 1 /* tcgen05: fp16 operands (saturating conversions), fp32 accumulate in TMEM; fp32 residual stream */
#define FFD_PREC_BF16 FFD_PREC_FP16 /* former name: the operands were bf16 until the switch to fp16 (same speed, 8x the
                                       mantissa: csrc/ffd_tc_ptx.cuh FFD_TC_FP16) */

#define FFD_SDE_VP 0
#define FFD_SDE_VE 1

/* score-model call mode (cached_transformer.py:106-329 picks the branch from n_own) */
#define FFD_MODE_STOCK 0  /* cache is None: nn.TransformerEncoder */
#define FFD_MODE_CACHED 1 /* cached layers: branch on the recompute set */

typedef void* ffd_stream_t; /* cudaStream_t */

typedef struct ffd_dims {
  int32_t max_len;    /* L */
  int32_t n_channels; /* C */
  int32_t d_model;    /* d */
  int32_t n_head;     /* H  (d % H == 0) */
  int32_t d_ff;       /* F  (2048 in the reference: score_models.py:61-63) */
  int32_t n_layers;   /* NL */
} ffd_dims;

/* Flat fp32 weight arena, element offsets (see ffd_weights_layout):
 *   pos[L*d] emb_w[d*C] emb_b[d] unemb_w[C*d] unemb_b[C]
 *   per layer: in_w[3d*d] in_b[3d] out_w[d*d] out_b[d] w1[F*d] b1[F] w2[d*F] b2[d]
 *              n1w[d] n1b[d] n2w[d] n2b[d]
 * (state_dict order of score_models.py:55-66; q/k/v are rows 0:d/d:2d/2d:3d of in_w,
 *  cached_transformer.py:94-100).  `pos` must already be max-norm renormed. */
int64_t ffd_weights_count(const ffd_dims* dims);
int64_t ffd_layer_stride(const ffd_dims* dims);      /* floats per layer block */
int64_t ffd_layer0_offset(const ffd_dims* dims);     /* offset of layer 0's in_w */

/* bf16 tensor-core image of the arena (UMMA canonical K-major core-matrix tiles) */
int64_t ffd_packed_bytes(const ffd_dims* dims);
int ffd_pack_weights(const ffd_dims* dims, const float* weights, void* packed, ffd_stream_t stream);

/* workspace for a batch of B samples */
int64_t ffd_workspace_bytes(const ffd_dims* dims, int64_t B, int precision);

const char* ffd_last_error(void);
int ffd_version(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches) */
uint64_t ffd_launch_count(void);
/* 1 if the tcgen05 path was compiled in and the current device is sm_100 */
int ffd_has_tensor_path(void);

/* ---- boundary transforms (fourier.py) ---------------------------------- */
/* U[L,L] (inverse=0) or V=U^-1[L,L] (inverse=1), fp32, entries evaluated in fp64 */
int ffd_dft_matrix(float* out, int L, int inverse, ffd_stream_t stream);
/* y[b,k,c] = sum_n T[k,n] * (x[b,n,c]*scale[n,c] + shift[n,c]); scale/shift may be NULL
 * (cmd/sample.py:107-113 de-standardise + idft fused) */
int ffd_dft_apply(const float* T, const float* x, const float* scale, const float* shift, float* y,
                  int64_t B, int L, int C, ffd_stream_t stream);

/* FreSca score filter (src/fdiff/utils/fresca.py:111-268, called from sampler.py:79-93): rfft along L, low band
 * scaled by low_scale, high band by high_scale, irfft.  U / V = ffd_dft_matrix(L, 0 / 1).  strategy 0 = "spatial"
 * cut-off (k <= cutoff_ratio * n_freq), 1 = "energy" (first bin whose cumulative batch-mean magnitude reaches
 * cutoff_ratio of the total; decided on the device, no host sync).  scratch: ffd_fresca_scratch_floats floats. */
int64_t ffd_fresca_scratch_floats(int64_t B, int L, int C);
int ffd_fresca(const float* U, const float* V, const float* score, float* out, float* scratch, int64_t B, int L, int C,
               float low_scale, float high_scale, float cutoff_ratio, int strategy, ffd_stream_t stream);

/* ---- E2-CRF event intensity (caching.py:524-559, E2CRFCache.compute_event_intensity) ------------------
 * out[0] = mean over the `rows` = num_layers * max_len rows of || crf[row, :] - prev_crf[row, :] ||_2,
 * out[1] = min(1, out[0] / tau_0).  out is DEVICE memory: the value can feed a device-side decision without a host
 * sync (the reference calls .item()).  Deterministic (fixed reduction order). */
int ffd_event_intensity(const float* crf, const float* prev_crf, int64_t rows, int d, float tau_0, float* out,
                        ffd_stream_t stream);

/* ---- sample-quality metrics (sampling/metrics.py:101-216 over utils/wasserstein.py:96-199) ---------------
 * Sliced / marginal Wasserstein-2 distances without leaving the device.  The reference projects with numpy and calls
 * POT's ot.emd2_1d (uniform weights, squared-Euclidean cost) per direction on the host.
 *   ffd_metric_project:   out[r, n] = <data[n, :], dirs[r, :]>            (wasserstein.py:146-149, all directions)
 *   ffd_metric_sort_rows: every row of x[R, N] ascending, in place         (N <= 32768)
 *   ffd_metric_w2_sorted: out[r] = sqrt(emd2_1d(a[r, :N], b[r, :M])) for ascending rows (wasserstein.py:117-143)
 *   ffd_metric_row_std:   out[r] = np.std(x[r, :])                         (wasserstein.py:157-159 "standardise")
 *   ffd_spectral_density: out[b, k, c] = |X_k|^2 of the real-packed spectrum x_tilde[B, L, C] -> [B, L/2+1, C]
 *                         (fourier.py:97-131 with apply_dft=False; MetricCollection's spectral metric) */
int ffd_spectral_density(const float* x_tilde, float* out, int64_t B, int L, int C, ffd_stream_t stream);
int ffd_metric_project(const float* data, int64_t N, int D, const float* dirs, int R, float* out, ffd_stream_t stream);
int ffd_metric_sort_rows(float* x, int R, int64_t N, ffd_stream_t stream);
int ffd_metric_w2_sorted(const float* a, int64_t N, const float* b, int64_t M, int R, float* out, ffd_stream_t stream);
int ffd_metric_row_std(const float* x, int R, int64_t N, float* out, ffd_stream_t stream);

/* ---- embeddings (transformer.py) ---------------------------------------- */
int ffd_posemb_renorm(float* table, int L, int d, float max_norm, ffd_stream_t stream);
/* out[i, :] = dense_w @ [sin(2*pi*t_i*W), cos(...)][:d] + dense_b   for i < n */
int ffd_time_embedding(const float* t, int n, const float* W, const float* dense_w, const float* dense_b,
                       int d, float* out, ffd_stream_t stream);

/* ---- score model --------------------------------------------------------- */
typedef struct ffd_forward_args {
  ffd_dims dims;
  int32_t precision;      /* FFD_PREC_* */
  int32_t mode;           /* FFD_MODE_* */
  int64_t B;
  const float* weights;   /* fp32 arena */
  const void* packed;     /* 16-bit operand images (FFD_PREC_FP16 only) */
  void* workspace;
  int64_t workspace_bytes;
  const float* X;         /* [B, L, C] */
  const float* time_emb;  /* [d]  (one shared t per batch: sampler.py:59-60) */
  /* cached mode */
  int32_t n_own;          /* |S| */
  const int32_t* own_mask;/* [L] 1 = token in S; may be NULL when n_own is 0 or L */
  float* k_cache;         /* [NL, H, L, hd] sample-0 K  (caching.py:364-377) */
  float* v_cache;
  float* crf_out;         /* [NL, L, d] or NULL  (score_models.py:151,188) */
  float* score;           /* [B, L, C] out; may be NULL if fused_step != NULL */
} ffd_forward_args;

int ffd_score_forward(const ffd_forward_args* a, ffd_stream_t stream);

/* y = LayerNorm2(x + linear2(relu(linear1(x)))) of one encoder layer on M token rows
 * (cached_transformer.py:130-132,325-327): the kernel that carries 86 % of the path's FLOPs.
 * x, y: [M, d] fp32 (y may alias x); scratch: [M, d] fp32, used by FFD_PREC_FP32 only. */
int ffd_ffn_ln(const ffd_dims* dims, int precision, const float* weights, const void* packed, int layer,
               const float* x, float* y, float* scratch, int64_t M, ffd_stream_t stream);

/* y = LayerNorm1(x + out_proj(softmax(q K_c^T / sqrt(hd)) V_c)) of one PURE-CACHE encoder layer on M token rows
 * (cached_transformer.py:228-258,309-321): K_c/V_c are the sample-0 cache rows [NL, H, L, hd] shared by the batch.
 * x, y: [M, d] fp32 (y may alias x for FFD_PREC_FP16); workspace as for ffd_score_forward with B = ceil(M / L). */
int ffd_attn_cached_ln(const ffd_dims* dims, int precision, const float* weights, const void* packed, int layer,
                       const float* k_cache, const float* v_cache, const float* x, float* y, void* workspace,
                       int64_t workspace_bytes, int64_t M, ffd_stream_t stream);

/* One whole PURE-CACHE encoder layer, y = LayerNorm2(y1 + FFN(y1)) with y1 the result of ffd_attn_cached_ln
 * (cached_transformer.py:228-258,309-327): the two blocks back to back.  With FFD_FUSED_LAYER=1 (opt-in, default model
 * width only) FFD_PREC_FP16 runs it as ONE kernel that overlaps the MUFU-bound softmax with the tensor-bound FFN GEMMs
 * per SM -- correct, but hand-off-latency bound and slower than the two kernels on B200, hence
 * not the default.  x, y: [M, d] fp32 (y may alias x for FFD_PREC_FP16). */
int ffd_layer_cached_ln(const ffd_dims* dims, int precision, const float* weights, const void* packed, int layer,
                        const float* k_cache, const float* v_cache, const float* x, float* y, void* workspace,
                        int64_t workspace_bytes, int64_t M, ffd_stream_t stream);

/* ---- SDE ------------------------------------------------------------------ */
typedef struct ffd_step_coef {
  int32_t kind;     /* FFD_SDE_* */
  float c_x;        /* VP: fp32(-0.5*beta)   VE: unused */
  float coef;       /* VP: fp32(sqrt(beta))  VE: fp32(sqrt_derivative) */
  float dt;         /* step_size (sde.py:64) */
  float sqrt_dt;    /* fp32 sqrt(dt) */
} ffd_step_coef;

/* x_out = x - drift*dt + sqrt_dt*(coef*G)*z, same rounding order as sde.py:228-244 / :143-163.
 * z == NULL: z is drawn in-kernel, Philox-4x32-10 keyed (seed, stream_id), element index = counter. */
int ffd_sde_step(const ffd_step_coef* c, const float* G, const float* x, const float* score, const float* z,
                 uint64_t seed, uint64_t stream_id, float* x_out, int64_t B, int L, int C, ffd_stream_t stream);
int ffd_prior(const float* G, const float* z, uint64_t seed, uint64_t stream_id, float scale, float* x_out,
              int64_t B, int L, int C, ffd_stream_t stream);
int ffd_philox_normal(uint64_t seed, uint64_t stream_id, float* out, int64_t n, ffd_stream_t stream);

/* ---- E2-CRF event-driven schedule, decided on the device (opt-in) ---------------------------------------
 * The reference's shipped schedule is a pure function of the step index (caching.py:131-181); its event code
 * (caching.py:524-559) is never called.  This is the schedule the reference's paper describes
 * (submission/TSD_Cache_Paper_Final.pdf, section 3.3, Eq. 14-18, Algorithm 1), run without a host sync:
 *
 *   after the score model of step i (ffd_event_decide; z = final-layer CRF of sample 0, [L, d]):
 *     r_i      = ||z_i - z_{i-1}||_F^2 / (||z_{i-1}||_F^2 + eta)                                    (Eq. 14)
 *     delta_k  = ||z_i[k] - z_{i-1}[k]||_2 ;  a_k <- a_k + delta_k   (error-feedback accumulator; a_k <- delta_k
 *                when accumulate == 0, which is Eq. 15 verbatim)
 *     tau_k    = tau_0 / (eps + ||x_i[k]||_2^2)            (energy-weighted threshold, x = sample 0's state)
 *     S_{i+1}  = {k <= K}  U  {k : a_k > tau_k}  U  RandomProbe(k > K, probability probe_ratio)      (Eq. 15)
 *     weight_k = 1 for k in S_{i+1} (recompute; a_k <- 0), else gamma_i = min(0.1, r_i / 2) when
 *                (i + 1) % R == 0 or r_i > tau_warn (error-feedback correction, Eq. 17-18), else 0 (reuse)
 *   before every layer's attention of step i+1 (ffd_event_refresh_kv):
 *     K_l[k], V_l[k] <- (1 - weight_k) * cache + weight_k * Project(layer input of sample 0, token k)
 *   and every sample attends to the refreshed sample-0 cache (pure-cache branch, cached_transformer.py:228-258).
 * With batch size 1 (the reference's benchmark batch, cmd/benchmark_cache.py:73) this is Algorithm 1 itself; for
 * larger batches the cache keeps sample 0 only, as the reference's store does (caching.py:326-328).
 * All buffers are DEVICE memory; the decision never leaves the device. */
#define FFD_EVENT_STATS 8 /* floats: [0] steps decided, [1] tokens recomputed (weight 1), [2] tokens corrected
                             (0 < weight < 1), [3] feedback steps, [4] last r_i, [5] sum of r_i, [6..7] reserved */
typedef struct ffd_event_args {
  float* z_prev;   /* [L, d]  final-layer CRF of the previous step */
  float* accum;    /* [L]     error-feedback accumulators a_k */
  float* weight;   /* [L]     blend weights applied by the NEXT step's refresh */
  float* stats;    /* [FFD_EVENT_STATS] running counters */
  int32_t K;       /* tokens 0..K are always recomputed */
  int32_t R;       /* error-feedback period (> 0) */
  int32_t accumulate; /* 1: a_k accumulates while a token is reused; 0: a_k = delta_k */
  float tau_0, tau_warn, eps, eta, probe_ratio;
  uint64_t seed;   /* RandomProbe: Philox keyed (seed, step), counter = token */
} ffd_event_args;

/* first != 0: (re)initialise -- z_prev <- z, a_k = 0, weight = 0, stats = 0 (after the full-recompute step 0). */
int ffd_event_decide(const ffd_event_args* ev, const float* z, const float* x0, int L, int d, int C, int step,
                     int first, ffd_stream_t stream);
/* layer `layer` of the sample-0 cache [NL, H, L, hd] from h0 = sample 0's rows [L, d] of that layer's input */
int ffd_event_refresh_kv(const ffd_dims* dims, const float* weights, int layer, const float* h0, const float* weight,
                         float* k_cache, float* v_cache, ffd_stream_t stream);

/* ---- the step loop -------------------------------------------------------- */
/* FreSca between the score model and the SDE update of every step (sampler.py:79-93), inside the fused loop: no host
 * synchronisation, both cut-off strategies decided on the device. */
typedef struct ffd_fresca_args {
  const float* U;        /* ffd_dft_matrix(L, 0) */
  const float* V;        /* ffd_dft_matrix(L, 1) */
  float* scratch;        /* DEVICE, ffd_fresca_scratch_floats(B, L, C) + 2 * B * L * C floats (raw + filtered score) */
  const float* high_scale_host; /* [n_steps] HOST array: the time-dependent high-band scale (fresca.py:243-256) */
  float low_scale;
  float cutoff_ratio;
  int32_t strategy;      /* 0 spatial, 1 energy */
} ffd_fresca_args;

typedef struct ffd_sample_args {
  ffd_forward_args fwd;        /* X is updated in place; score may be NULL */
  const float* G;              /* [L] */
  const float* time_emb_table; /* [n_steps, d] */
  const ffd_step_coef* coef_host;   /* [n_steps] HOST array */
  const int32_t* n_own_host;   /* [n_steps] HOST array, -1 = no cache (stock) */
  const int32_t* store_crf_host; /* [n_steps] HOST array: 1 = write crf_out this step (caching.py:474) */
  const float* noise;          /* injected z for steps noise_step0.., [*, B, L, C]; NULL = Philox in-kernel */
  int32_t noise_step0;         /* absolute step index of noise[0] */
  uint64_t seed;
  uint64_t stream_base;        /* Philox stream id of step 0 */
  float* x;                    /* [B, L, C] state, in/out */
  float* x_tmp;                /* [B, L, C] scratch */
  int32_t step_begin, step_end;
  const ffd_event_args* event; /* HOST struct or NULL: device-decided schedule for cached steps with n_own_host == 0 */
  const ffd_fresca_args* fresca; /* HOST struct or NULL: filter the score of every step (a step whose scales are both
                                    1 is left alone, fresca.py:150-152) */
  int32_t graph;               /* != 0: runs of pure-cache steps are captured once (two steps) as a CUDA graph and replayed;
                                  per-step scalars are then read from graph_scratch on the device.  For the launch-bound
                                  regime (small batches): same kernels, same results. */
  void* graph_scratch;         /* DEVICE, ffd_graph_scratch_bytes(step_end) bytes (only with graph != 0) */
  int64_t graph_scratch_bytes;
} ffd_sample_args;

int64_t ffd_graph_scratch_bytes(int32_t n_steps);

int ffd_sample_steps(const ffd_sample_args* a, ffd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FFD_B200_H */
