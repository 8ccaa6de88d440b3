// C ABI (include/ffd_b200.h) and the per-step orchestration of the score model.
#include <stdarg.h>
#include <string.h>

#include "ffd_internal.h"

namespace ffd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }

// every branch of this model runs on the tensor-core kernels: no fp32 staging buffers are needed (carve)
static bool lean_workspace(const ffd_dims& m, int precision) {
  return precision == FFD_PREC_FP16 && tc_attn_supported(m) && tc_attn_own_fused(m);
}

static int check_dims(const ffd_dims& m) {
  FFD_CHECK_ARG(m.max_len > 0 && m.n_channels > 0 && m.d_model > 0 && m.n_head > 0 && m.d_ff > 0 && m.n_layers > 0,
                "non-positive model dimension");
  FFD_CHECK_ARG(m.d_model % m.n_head == 0, "d_model %d not divisible by n_head %d", m.d_model, m.n_head);
  FFD_CHECK_ARG(m.d_model <= 128, "d_model %d > 128 unsupported", m.d_model);
  FFD_CHECK_ARG(m.d_model / m.n_head <= 64, "head_dim %d > 64 unsupported", m.d_model / m.n_head);
  return FFD_OK;
}

// one encoder layer, fp32 arithmetic (FFD_PREC_FP16 swaps in the tcgen05 FFN)
//   x = LN1(x + MHA(x));  x = LN2(x + FFN(x))      cached_transformer.py:123-134 / :309-327
static int run_layer(const ffd_forward_args& a, const Arena& ar, const Workspace& ws, int layer, int branch,
                     cudaStream_t s) {
  const ffd_dims& m = a.dims;
  const int L = m.max_len, d = m.d_model, H = m.n_head, hd = d / H, F = m.d_ff;
  const int64_t M = a.B * L;
  const float* w = a.weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  float* kc = a.k_cache ? a.k_cache + (int64_t)layer * H * L * hd : nullptr;
  float* vc = a.v_cache ? a.v_cache + (int64_t)layer * H * L * hd : nullptr;
  int rc;
  // branch: 0 = stock attention (all K/V own), 1 = pure cache, 2 = mixed
  if (branch == 1 && a.precision == FFD_PREC_FP16 && tc_layer_enabled(m)) {
    // pure-cache layer on tensor cores, ONE launch: attention block and FFN block overlapped on every SM
    return tc_layer_cached(m, a.packed, ws.kvimg96, a.weights, layer, ws.h, ws.h, M, s);
  }
  if (branch == 1 && a.precision == FFD_PREC_FP16 && tc_attn_supported(m)) {
    // the same as two launches (models the fused kernel is not instantiated for, or FFD_FUSED_LAYER=0)
    if ((rc = tc_attn_cached(m, tc_attn_image(m, a.packed, layer), ws.kvimg, a.weights, layer, ws.h, ws.h, M, s)))
      return rc;
    return tc_ffn_ln(m, a.packed, a.weights, layer, ws.h, ws.h, M, s);
  }
  if ((branch == 0 || branch == 2) && a.precision == FFD_PREC_FP16 && tc_attn_supported(m)) {
    // stock attention (all K/V own) on the same tensor-core kernel: K/V projection + per-sample operand images
    // (one tcgen05 kernel; fp32 projection + pack kernel for head dims > 7), attention block, FFN block.  Q is
    // projected inside the attention kernel.  Mixed branch (2): the same, with the key rows outside the recompute
    // set patched in from the shared sample-0 images, and sample 0's recomputed rows (projected in fp32 from the
    // layer INPUT, cached_transformer.py:282-304) stored afterwards.
    if (branch == 2 && (rc = launch_linear(ws.h, d, w + ar.in_w + (int64_t)d * d, d, w + ar.in_b + d, ws.kvtmp, 2 * d,
                                           L, 2 * d, d, 0, s)))
      return rc;
    const float* kv_src = nullptr;
    if (!tc_attn_own_fused(m)) {
      if ((rc = launch_linear(ws.h, d, w + ar.in_w + (int64_t)d * d, d, w + ar.in_b + d, ws.qkv + d, 3 * d, M, 2 * d,
                              d, 0, s)))
        return rc;
      kv_src = ws.qkv;
    }
    // the per-sample operand images are built and consumed OWN_CHUNK samples at a time
    for (int64_t b0 = 0; b0 < a.B; b0 += OWN_CHUNK) {
      const int64_t nb = a.B - b0 < OWN_CHUNK ? a.B - b0 : OWN_CHUNK;
      float* hc = ws.h + b0 * L * d;
      if ((rc = tc_attn_own(m, tc_attn_image(m, a.packed, layer), ws.kvimg_own, kv_src ? kv_src + b0 * L * 3 * d : nullptr,
                            3 * d, d, 2 * d, a.weights, layer, hc, hc, nb * L, s, branch == 2 ? ws.kvimg : nullptr,
                            a.n_own, a.own_mask)))
        return rc;
    }
    if (branch == 2 && (rc = launch_scatter_kv(ws.kvtmp, 2 * d, 0, d, a.own_mask, a.n_own, kc, vc, L, H, hd, s)))
      return rc;
    return tc_ffn_ln(m, a.packed, a.weights, layer, ws.h, ws.h, M, s);
  }
  const int nq = (branch == 1) ? d : 3 * d;  // pure cache projects Q only (cached_transformer.py:229)
  if ((rc = launch_linear(ws.h, d, w + ar.in_w, d, w + ar.in_b, ws.qkv, 3 * d, M, nq, d, 0, s))) return rc;
  if (branch == 2) {
    // store sample 0's recomputed rows BEFORE attention reads the cache? No: the mixed branch
    // reads cached rows only for tokens outside S and writes only rows inside S
    // (cached_transformer.py:275-276,295-296,304), so the order is immaterial.
    if ((rc = launch_scatter_kv(ws.qkv, 3 * d, d, 2 * d, a.own_mask, a.n_own, kc, vc, L, H, hd, s))) return rc;
  }
  const int n_prefix = (branch == 0) ? L : (branch == 1 ? 0 : a.n_own);
  const int32_t* mask = (branch == 2) ? a.own_mask : nullptr;
  if ((rc = launch_attention(ws.qkv, 3 * d, kc, vc, mask, n_prefix, ws.attn, a.B, L, d, H, s))) return rc;
  if ((rc = launch_linear(ws.attn, d, w + ar.out_w, d, w + ar.out_b, ws.tmp, d, M, d, d, 0, s))) return rc;
  if ((rc = launch_add_ln(ws.h, ws.tmp, w + ar.n1w, w + ar.n1b, ws.h, M, d, s))) return rc;
  if (a.precision == FFD_PREC_FP16) {
    if ((rc = tc_ffn_ln(m, a.packed, a.weights, layer, ws.h, ws.h, M, s))) return rc;
  } else {
    if ((rc = launch_ffn(ws.h, w + ar.w1, w + ar.b1, w + ar.w2, w + ar.b2, ws.tmp, M, d, F, s))) return rc;
    if ((rc = launch_add_ln(ws.h, ws.tmp, w + ar.n2w, w + ar.n2b, ws.h, M, d, s))) return rc;
  }
  return FFD_OK;
}

// Which branch of CachedTransformerEncoderLayer.forward runs (cached_transformer.py:106-329).
//   returns 0 stock, 1 pure cache, 2 mixed; *populate = 1 when the full-recompute branch stores K/V.
static int pick_branch(const ffd_forward_args& a, int* populate) {
  *populate = 0;
  if (a.mode == FFD_MODE_STOCK) return 0;
  const int L = a.dims.max_len;
  if (a.n_own == L) {
    *populate = 1;
    return 0;
  }
  if ((double)a.n_own > 0.8 * (double)L) return 0;
  return a.n_own == 0 ? 1 : 2;
}

// ev != nullptr: the device-decided schedule (ffd_event_args) -- pure-cache steps refresh each layer's sample-0 K/V
// rows under the weights the previous step's decision left on the device, and every cached step ends with the
// next decision; the full-recompute step (re)initialises the event state.
static int forward_hidden(const ffd_forward_args& a, const Arena& ar, const Workspace& ws, cudaStream_t s,
                          const ffd_event_args* ev = nullptr, int step = 0, const int32_t* step_ctr = nullptr) {
  const ffd_dims& m = a.dims;
  const int L = m.max_len, C = m.n_channels, d = m.d_model, H = m.n_head, hd = d / H;
  const int64_t M = a.B * L;
  int rc;
  if ((rc = launch_embed(a.X, a.weights + ar.emb_w, a.weights + ar.emb_b, a.weights + ar.pos, a.time_emb, ws.h, M,
                         L, C, d, s, step_ctr)))
    return rc;
  int populate = 0;
  const int branch = pick_branch(a, &populate);
  const bool tc_images = (branch == 1 || branch == 2) && a.precision == FFD_PREC_FP16 && tc_attn_supported(m);
  const bool event_refresh = ev && branch == 1;
  if (tc_images && !event_refresh) {
    // the fp32 cache is the source of truth (it may have been reset or refreshed since the last
    // call); its bf16 operand images are 1.5 MB, so they are simply rebuilt every pure-cache step
    const bool fused = branch == 1 && tc_layer_enabled(m);
    if (fused ? (rc = tc_pack_kv96(m, a.k_cache, a.v_cache, ws.kvimg96, s))
              : (rc = tc_pack_kv(m, a.k_cache, a.v_cache, ws.kvimg, s)))
      return rc;
  }
  for (int l = 0; l < m.n_layers; ++l) {
    if (event_refresh) {
      // sample 0's rows of this layer's input are the first L rows of ws.h
      const float* w = a.weights + ar.layer0 + (int64_t)l * ar.layer_stride;
      if ((rc = launch_event_refresh_kv(w + ar.in_w + (int64_t)d * d, w + ar.in_b + d, ws.h, ev->weight,
                                        a.k_cache + (int64_t)l * H * L * hd, a.v_cache + (int64_t)l * H * L * hd, L, d,
                                        H, s)))
        return rc;
      if (tc_images && (tc_layer_enabled(m) ? (rc = tc_pack_kv96_layer(m, l, a.k_cache, a.v_cache, ws.kvimg96, s))
                                            : (rc = tc_pack_kv_layer(m, l, a.k_cache, a.v_cache, ws.kvimg, s))))
        return rc;
    }
    if ((rc = run_layer(a, ar, ws, l, branch, s))) return rc;
    if (populate) {
      // K/V of the layer OUTPUT of sample 0, all rows (cached_transformer.py:142-168)
      const float* w = a.weights + ar.layer0 + (int64_t)l * ar.layer_stride;
      if ((rc = launch_linear(ws.h, d, w + ar.in_w + (int64_t)d * d, d, w + ar.in_b + d, ws.kvtmp, 2 * d, L, 2 * d,
                              d, 0, s)))
        return rc;
      if ((rc = launch_scatter_kv(ws.kvtmp, 2 * d, 0, d, nullptr, L, a.k_cache + (int64_t)l * H * L * hd,
                                  a.v_cache + (int64_t)l * H * L * hd, L, H, hd, s)))
        return rc;
    }
    if (a.crf_out && a.mode == FFD_MODE_CACHED) {
      // crf[l] = h[0]  (score_models.py:151,188)
      FFD_CHECK_CUDA(cudaMemcpyAsync(a.crf_out + (int64_t)l * L * d, ws.h, sizeof(float) * L * d,
                                     cudaMemcpyDeviceToDevice, s));
    }
  }
  if (ev && a.mode == FFD_MODE_CACHED && (event_refresh || populate)) {
    // z = final-layer CRF of sample 0, x = sample 0's state: the decision for the next step stays on the device
    if ((rc = launch_event_decide(*ev, ws.h, a.X, L, d, C, step, populate ? 1 : 0, s))) return rc;
  }
  return FFD_OK;
}

static int check_forward(const ffd_forward_args& a) {
  int rc = check_dims(a.dims);
  if (rc) return rc;
  FFD_CHECK_ARG(a.B >= 0, "negative batch");
  FFD_CHECK_ARG(a.weights && a.workspace && a.X && a.time_emb, "null pointer in ffd_forward_args");
  FFD_CHECK_ARG(a.precision == FFD_PREC_FP32 || a.precision == FFD_PREC_FP16, "unknown precision %d", a.precision);
  if (a.precision == FFD_PREC_FP16) {
    FFD_CHECK_ARG(a.packed != nullptr, "FFD_PREC_FP16 needs the packed weight image");
    if (!tc_supported(a.dims)) {
      set_error("tensor-core path unsupported for these dims (needs d_model <= 80, d_ff %% 128 == 0)");
      return FFD_ERR_UNSUPPORTED;
    }
  }
  FFD_CHECK_ARG(a.mode == FFD_MODE_STOCK || a.mode == FFD_MODE_CACHED, "unknown mode %d", a.mode);
  if (a.mode == FFD_MODE_CACHED) {
    FFD_CHECK_ARG(a.k_cache && a.v_cache, "cached mode needs k_cache/v_cache");
    FFD_CHECK_ARG(a.n_own >= 0 && a.n_own <= a.dims.max_len, "n_own %d out of range", a.n_own);
  }
  const int64_t need = ffd_workspace_bytes(&a.dims, a.B, a.precision);
  if (a.workspace_bytes < need) {
    set_error("workspace too small: %lld < %lld", (long long)a.workspace_bytes, (long long)need);
    return FFD_ERR_WORKSPACE;
  }
  return FFD_OK;
}

}  // namespace ffd


// ---- the step loop as a replayed CUDA graph (launch-bound regime: small batches) ---------------------------------
// A run of consecutive PURE-CACHE steps (no CRF store, no event schedule, no FreSca) launches the same ~23 kernels per
// step with only four things changing: the time-embedding row, the SDE scalars, the Philox stream and the slice of
// injected noise.  Those are read from device memory through a step cursor (StepSource, ffd_common.cuh), so TWO steps
// (the state ping-pongs between x and x_tmp) are captured once and the graph is replayed: one launch per two steps
// instead of 46.  At batch 64 the loop is launch-bound (0.54 ms per step for ~0.2 ms of GPU work).
namespace ffd {
struct GraphKey {
  ffd_dims dims;
  int32_t precision;
  int64_t B;
  const void *weights, *packed, *workspace, *k_cache, *v_cache, *score, *G, *te, *noise, *x, *x_tmp, *scratch;
  int fused;
};
struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t exec;
  unsigned long long launches;  // kernels per replay
  bool used;
};
static GraphEntry g_graphs[8];
static int g_graph_next = 0;
static bool g_pure_step_ran = false;  // a pure-cache step has run directly: every lazy initialisation of the launchers is done

static size_t graph_coef_offset() { return 32; }
}  // namespace ffd

namespace ffd {
extern long long* g_attn_dbg;
void tc_debug_set_ffn_pairs(int mode);
void tc_debug_set_attn_quarters(int mode);
}
using namespace ffd;

extern "C" {

void ffd_debug_set_attn_stamps(long long* dev_buf) { ffd::g_attn_dbg = dev_buf; }
// test / A-B hook (not part of the ABI header): 1 = FFN on CTA pairs, 0 = one-CTA kernel, -1 = FFD_FFN_CTA2 / default
void ffd_debug_set_ffn_pairs(int mode) { ffd::tc_debug_set_ffn_pairs(mode); }
// 1 = attention kernel v12 (four independent quarter streams), 0 = v11, -1 = FFD_ATTN_QUARTERS / default
void ffd_debug_set_attn_quarters(int mode) { ffd::tc_debug_set_attn_quarters(mode); }
// 1 = pure-cache layers as ONE fused kernel, 0 = attention + FFN kernels, -1 = FFD_FUSED_LAYER / default (off)
void ffd_debug_set_fused_layer(int mode) { ffd::tc_debug_set_fused_layer(mode); }
int ffd_version(void) { return 1; }
uint64_t ffd_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
const char* ffd_last_error(void) { return g_err; }

int64_t ffd_weights_count(const ffd_dims* dims) { return dims ? arena_of(*dims).total : -1; }
int64_t ffd_layer_stride(const ffd_dims* dims) { return dims ? arena_of(*dims).layer_stride : -1; }
int64_t ffd_layer0_offset(const ffd_dims* dims) { return dims ? arena_of(*dims).layer0 : -1; }

int64_t ffd_packed_bytes(const ffd_dims* dims) { return dims ? tc_packed_bytes(*dims) : -1; }

int ffd_pack_weights(const ffd_dims* dims, const float* weights, void* packed, ffd_stream_t stream) {
  FFD_CHECK_ARG(dims && weights && packed, "null pointer");
  int rc = check_dims(*dims);
  if (rc) return rc;
  if (!tc_supported(*dims)) {
    set_error("tensor-core path unsupported for these dims");
    return FFD_ERR_UNSUPPORTED;
  }
  return tc_pack_weights(*dims, weights, packed, (cudaStream_t)stream);
}

int64_t ffd_workspace_bytes(const ffd_dims* dims, int64_t B, int precision) {
  if (!dims || B < 0) return -1;
  return carve(*dims, B, nullptr, lean_workspace(*dims, precision)).bytes;
}

int ffd_has_tensor_path(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int ffd_dft_matrix(float* out, int L, int inverse, ffd_stream_t stream) {
  FFD_CHECK_ARG(out && L > 0 && L <= 8192, "bad ffd_dft_matrix arguments");
  return launch_dft_matrix(out, L, inverse, (cudaStream_t)stream);
}

int ffd_dft_apply(const float* T, const float* x, const float* scale, const float* shift, float* y, int64_t B,
                  int L, int C, ffd_stream_t stream) {
  FFD_CHECK_ARG(T && x && y && B >= 0 && L > 0 && C > 0, "bad ffd_dft_apply arguments");
  FFD_CHECK_ARG(x != y, "ffd_dft_apply cannot run in place");
  return launch_dft_apply(T, x, scale, shift, y, B, L, C, (cudaStream_t)stream);
}

int64_t ffd_fresca_scratch_floats(int64_t B, int L, int C) { return fresca_scratch_floats(B, L, C); }

int ffd_fresca(const float* U, const float* V, const float* score, float* out, float* scratch, int64_t B, int L, int C,
               float low_scale, float high_scale, float cutoff_ratio, int strategy, ffd_stream_t stream) {
  FFD_CHECK_ARG(U && V && score && out && scratch && B >= 0 && L > 0 && C > 0, "bad ffd_fresca arguments");
  FFD_CHECK_ARG(strategy == 0 || strategy == 1, "unknown cutoff strategy %d (0 spatial, 1 energy)", strategy);
  FFD_CHECK_ARG(score != out, "ffd_fresca cannot run in place");
  if (B == 0) return FFD_OK;
  return launch_fresca(U, V, score, out, scratch, B, L, C, low_scale, high_scale, cutoff_ratio, strategy,
                       (cudaStream_t)stream);
}

int ffd_event_intensity(const float* crf, const float* prev_crf, int64_t rows, int d, float tau_0, float* out,
                        ffd_stream_t stream) {
  FFD_CHECK_ARG(crf && prev_crf && out && rows > 0 && d > 0, "bad ffd_event_intensity arguments");
  FFD_CHECK_ARG(tau_0 > 0.f, "tau_0 must be positive");
  return launch_event_intensity(crf, prev_crf, rows, d, tau_0, out, (cudaStream_t)stream);
}

int ffd_spectral_density(const float* x_tilde, float* out, int64_t B, int L, int C, ffd_stream_t stream) {
  FFD_CHECK_ARG(x_tilde && out && B >= 0 && L > 0 && C > 0, "bad ffd_spectral_density arguments");
  if (B == 0) return FFD_OK;
  return launch_spectral_density(x_tilde, out, B, L, C, (cudaStream_t)stream);
}

int ffd_metric_project(const float* data, int64_t N, int D, const float* dirs, int R, float* out,
                       ffd_stream_t stream) {
  FFD_CHECK_ARG(data && dirs && out && N > 0 && D > 0 && R > 0, "bad ffd_metric_project arguments");
  return launch_metric_project(data, N, D, dirs, R, out, (cudaStream_t)stream);
}

int ffd_metric_sort_rows(float* x, int R, int64_t N, ffd_stream_t stream) {
  FFD_CHECK_ARG(x && R > 0 && N > 0, "bad ffd_metric_sort_rows arguments");
  if (N > 32768) {
    set_error("ffd_metric_sort_rows: rows of more than 32768 values are not supported (got %lld)", (long long)N);
    return FFD_ERR_UNSUPPORTED;
  }
  return launch_metric_sort_rows(x, R, N, (cudaStream_t)stream);
}

int ffd_metric_w2_sorted(const float* a, int64_t N, const float* b, int64_t M, int R, float* out,
                         ffd_stream_t stream) {
  FFD_CHECK_ARG(a && b && out && N > 0 && M > 0 && R > 0, "bad ffd_metric_w2_sorted arguments");
  return launch_metric_w2_sorted(a, N, b, M, R, out, (cudaStream_t)stream);
}

int ffd_metric_row_std(const float* x, int R, int64_t N, float* out, ffd_stream_t stream) {
  FFD_CHECK_ARG(x && out && N > 0 && R > 0, "bad ffd_metric_row_std arguments");
  return launch_metric_row_std(x, R, N, out, (cudaStream_t)stream);
}

int ffd_posemb_renorm(float* table, int L, int d, float max_norm, ffd_stream_t stream) {
  FFD_CHECK_ARG(table && L > 0 && d > 0, "bad ffd_posemb_renorm arguments");
  return launch_posemb_renorm(table, L, d, max_norm, (cudaStream_t)stream);
}

int ffd_time_embedding(const float* t, int n, const float* W, const float* dense_w, const float* dense_b, int d,
                       float* out, ffd_stream_t stream) {
  FFD_CHECK_ARG(t && W && dense_w && dense_b && out && n >= 0 && d > 0 && d <= 4096, "bad ffd_time_embedding arguments");
  return launch_time_embedding(t, n, W, dense_w, dense_b, d, out, (cudaStream_t)stream);
}

int ffd_score_forward(const ffd_forward_args* a, ffd_stream_t stream) {
  FFD_CHECK_ARG(a != nullptr, "null args");
  int rc = check_forward(*a);
  if (rc) return rc;
  FFD_CHECK_ARG(a->score != nullptr, "score output is null");
  if (a->B == 0) return FFD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const Arena ar = arena_of(a->dims);
  const Workspace ws = carve(a->dims, a->B, a->workspace, lean_workspace(a->dims, a->precision));
  if ((rc = forward_hidden(*a, ar, ws, s))) return rc;
  const ffd_dims& m = a->dims;
  return launch_unembed_step(ws.h, a->weights + ar.unemb_w, a->weights + ar.unemb_b, a->score, 0, nullptr, nullptr,
                             nullptr, nullptr, 0, 0, nullptr, a->B * m.max_len, m.max_len, m.n_channels, m.d_model,
                             s);
}

int ffd_ffn_ln(const ffd_dims* dims, int precision, const float* weights, const void* packed, int layer,
               const float* x, float* y, float* scratch, int64_t M, ffd_stream_t stream) {
  FFD_CHECK_ARG(dims && weights && x && y && M >= 0, "bad ffd_ffn_ln arguments");
  int rc = check_dims(*dims);
  if (rc) return rc;
  FFD_CHECK_ARG(layer >= 0 && layer < dims->n_layers, "layer %d out of range", layer);
  if (M == 0) return FFD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == FFD_PREC_FP16) {
    FFD_CHECK_ARG(packed != nullptr, "FFD_PREC_FP16 needs the packed weight image");
    if (!tc_supported(*dims)) {
      set_error("tensor-core path unsupported for these dims");
      return FFD_ERR_UNSUPPORTED;
    }
    return tc_ffn_ln(*dims, packed, weights, layer, x, y, M, s);
  }
  FFD_CHECK_ARG(precision == FFD_PREC_FP32, "unknown precision %d", precision);
  FFD_CHECK_ARG(scratch != nullptr, "fp32 path needs an [M, d] scratch buffer");
  const Arena ar = arena_of(*dims);
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  const int d = dims->d_model;
  if ((rc = launch_ffn(x, w + ar.w1, w + ar.b1, w + ar.w2, w + ar.b2, scratch, M, d, dims->d_ff, s))) return rc;
  return launch_add_ln(x, scratch, w + ar.n2w, w + ar.n2b, y, M, d, s);
}

int ffd_attn_cached_ln(const ffd_dims* dims, int precision, const float* weights, const void* packed, int layer,
                       const float* k_cache, const float* v_cache, const float* x, float* y, void* workspace,
                       int64_t workspace_bytes, int64_t M, ffd_stream_t stream) {
  FFD_CHECK_ARG(dims && weights && k_cache && v_cache && x && y && workspace && M >= 0, "bad ffd_attn_cached_ln arguments");
  int rc = check_dims(*dims);
  if (rc) return rc;
  const ffd_dims& m = *dims;
  FFD_CHECK_ARG(layer >= 0 && layer < m.n_layers, "layer %d out of range", layer);
  const int L = m.max_len, d = m.d_model, H = m.n_head, hd = d / H;
  FFD_CHECK_ARG(M % L == 0, "M must be a whole number of samples (M %% max_len == 0)");
  const int64_t B = M / L;
  if (workspace_bytes < ffd_workspace_bytes(dims, B, precision)) {
    set_error("workspace too small");
    return FFD_ERR_WORKSPACE;
  }
  if (M == 0) return FFD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const Arena ar = arena_of(m);
  const Workspace ws = carve(m, B, workspace, lean_workspace(m, precision));
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  if (precision == FFD_PREC_FP16) {
    FFD_CHECK_ARG(packed != nullptr, "FFD_PREC_FP16 needs the packed weight image");
    if (!tc_attn_supported(m)) {
      set_error("tensor-core attention unsupported for these dims");
      return FFD_ERR_UNSUPPORTED;
    }
    if ((rc = tc_pack_kv(m, k_cache, v_cache, ws.kvimg, s))) return rc;
    return tc_attn_cached(m, tc_attn_image(m, packed, layer), ws.kvimg, weights, layer, x, y, M, s);
  }
  FFD_CHECK_ARG(precision == FFD_PREC_FP32, "unknown precision %d", precision);
  const float* kc = k_cache + (int64_t)layer * H * L * hd;
  const float* vc = v_cache + (int64_t)layer * H * L * hd;
  if ((rc = launch_linear(x, d, w + ar.in_w, d, w + ar.in_b, ws.qkv, 3 * d, M, d, d, 0, s))) return rc;
  if ((rc = launch_attention(ws.qkv, 3 * d, kc, vc, nullptr, 0, ws.attn, B, L, d, H, s))) return rc;
  if ((rc = launch_linear(ws.attn, d, w + ar.out_w, d, w + ar.out_b, ws.tmp, d, M, d, d, 0, s))) return rc;
  return launch_add_ln(x, ws.tmp, w + ar.n1w, w + ar.n1b, y, M, d, s);
}

int ffd_layer_cached_ln(const ffd_dims* dims, int precision, const float* weights, const void* packed, int layer,
                        const float* k_cache, const float* v_cache, const float* x, float* y, void* workspace,
                        int64_t workspace_bytes, int64_t M, ffd_stream_t stream) {
  FFD_CHECK_ARG(dims && weights && k_cache && v_cache && x && y && workspace && M >= 0, "bad ffd_layer_cached_ln arguments");
  int rc = check_dims(*dims);
  if (rc) return rc;
  const ffd_dims& m = *dims;
  FFD_CHECK_ARG(layer >= 0 && layer < m.n_layers, "layer %d out of range", layer);
  FFD_CHECK_ARG(M % m.max_len == 0, "M must be a whole number of samples (M %% max_len == 0)");
  if (precision == FFD_PREC_FP16 && tc_layer_enabled(m)) {
    FFD_CHECK_ARG(packed != nullptr, "FFD_PREC_FP16 needs the packed weight image");
    const int64_t B = M / m.max_len;
    if (workspace_bytes < ffd_workspace_bytes(dims, B, precision)) {
      set_error("workspace too small");
      return FFD_ERR_WORKSPACE;
    }
    if (M == 0) return FFD_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const Workspace ws = carve(m, B, workspace, lean_workspace(m, precision));
    if ((rc = tc_pack_kv96(m, k_cache, v_cache, ws.kvimg96, s))) return rc;
    return tc_layer_cached(m, packed, ws.kvimg96, weights, layer, x, y, M, s);
  }
  // everything else: the attention block, then the FFN block (y doubles as the intermediate)
  if ((rc = ffd_attn_cached_ln(dims, precision, weights, packed, layer, k_cache, v_cache, x, y, workspace,
                               workspace_bytes, M, stream)))
    return rc;
  if (precision == FFD_PREC_FP16) return ffd_ffn_ln(dims, precision, weights, packed, layer, y, y, nullptr, M, stream);
  const Workspace ws = carve(m, M / m.max_len, workspace, false);
  return ffd_ffn_ln(dims, precision, weights, packed, layer, y, y, ws.tmp, M, stream);
}

int ffd_sde_step(const ffd_step_coef* c, const float* G, const float* x, const float* score, const float* z,
                 uint64_t seed, uint64_t stream_id, float* x_out, int64_t B, int L, int C, ffd_stream_t stream) {
  FFD_CHECK_ARG(c && G && x && score && x_out && B >= 0 && L > 0 && C > 0, "bad ffd_sde_step arguments");
  FFD_CHECK_ARG(c->kind == FFD_SDE_VP || c->kind == FFD_SDE_VE, "unknown SDE kind %d", c->kind);
  FFD_CHECK_ARG(c->dt > 0.f, "step_size must be positive (sde.py:158,239)");
  return launch_sde_step(c, G, x, score, z, seed, stream_id, x_out, B, L, C, (cudaStream_t)stream);
}

int ffd_prior(const float* G, const float* z, uint64_t seed, uint64_t stream_id, float scale, float* x_out,
              int64_t B, int L, int C, ffd_stream_t stream) {
  FFD_CHECK_ARG(G && x_out && B >= 0 && L > 0 && C > 0, "bad ffd_prior arguments");
  return launch_prior(G, z, seed, stream_id, scale, x_out, B, L, C, (cudaStream_t)stream);
}

int ffd_philox_normal(uint64_t seed, uint64_t stream_id, float* out, int64_t n, ffd_stream_t stream) {
  FFD_CHECK_ARG(out && n >= 0, "bad ffd_philox_normal arguments");
  return launch_philox_normal(seed, stream_id, out, n, (cudaStream_t)stream);
}

int ffd_event_decide(const ffd_event_args* ev, const float* z, const float* x0, int L, int d, int C, int step,
                     int first, ffd_stream_t stream) {
  FFD_CHECK_ARG(ev && z && x0 && L > 0 && d > 0 && C > 0 && step >= 0, "bad ffd_event_decide arguments");
  FFD_CHECK_ARG(ev->z_prev && ev->accum && ev->weight && ev->stats, "null pointer in ffd_event_args");
  FFD_CHECK_ARG(ev->R > 0 && ev->K >= 0, "ffd_event_args: R must be positive, K non-negative");
  return launch_event_decide(*ev, z, x0, L, d, C, step, first, (cudaStream_t)stream);
}

int ffd_event_refresh_kv(const ffd_dims* dims, const float* weights, int layer, const float* h0, const float* weight,
                         float* k_cache, float* v_cache, ffd_stream_t stream) {
  FFD_CHECK_ARG(dims && weights && h0 && weight && k_cache && v_cache, "bad ffd_event_refresh_kv arguments");
  int rc = check_dims(*dims);
  if (rc) return rc;
  const ffd_dims& m = *dims;
  FFD_CHECK_ARG(layer >= 0 && layer < m.n_layers, "layer %d out of range", layer);
  const Arena ar = arena_of(m);
  const int L = m.max_len, d = m.d_model, H = m.n_head, hd = d / H;
  const float* w = weights + ar.layer0 + (int64_t)layer * ar.layer_stride;
  return launch_event_refresh_kv(w + ar.in_w + (int64_t)d * d, w + ar.in_b + d, h0, weight,
                                 k_cache + (int64_t)layer * H * L * hd, v_cache + (int64_t)layer * H * L * hd, L, d, H,
                                 (cudaStream_t)stream);
}

int64_t ffd_graph_scratch_bytes(int32_t n_steps) {
  return n_steps < 0 ? -1 : (int64_t)graph_coef_offset() + (int64_t)n_steps * (int64_t)sizeof(ffd_step_coef) + 64;
}

int ffd_sample_steps(const ffd_sample_args* a, ffd_stream_t stream) {
  FFD_CHECK_ARG(a != nullptr, "null args");
  FFD_CHECK_ARG(a->G && a->time_emb_table && a->coef_host && a->n_own_host && a->x && a->x_tmp,
                "null pointer in ffd_sample_args");
  FFD_CHECK_ARG(a->step_begin >= 0 && a->step_end >= a->step_begin, "bad step range");
  FFD_CHECK_ARG(!a->noise || a->noise_step0 <= a->step_begin, "noise_step0 beyond step_begin");
  if (a->event) {
    const ffd_event_args& e = *a->event;
    FFD_CHECK_ARG(e.z_prev && e.accum && e.weight && e.stats, "null pointer in ffd_event_args");
    FFD_CHECK_ARG(e.R > 0 && e.K >= 0, "ffd_event_args: R must be positive, K non-negative");
    FFD_CHECK_ARG(a->fwd.k_cache && a->fwd.v_cache, "the event-driven schedule needs the K/V cache");
  }
  if (a->fresca) {
    const ffd_fresca_args& fr = *a->fresca;
    FFD_CHECK_ARG(fr.U && fr.V && fr.scratch && fr.high_scale_host, "null pointer in ffd_fresca_args");
    FFD_CHECK_ARG(fr.strategy == 0 || fr.strategy == 1, "unknown cutoff strategy %d (0 spatial, 1 energy)", fr.strategy);
  }
  ffd_forward_args f = a->fwd;
  f.X = a->x;
  f.time_emb = a->time_emb_table;  // placeholder for the null check
  int rc = check_forward(f);
  if (rc) return rc;
  if (f.B == 0) return FFD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const ffd_dims& m = f.dims;
  const Arena ar = arena_of(m);
  const Workspace ws = carve(m, f.B, f.workspace, lean_workspace(m, f.precision));
  const int64_t n_elem = f.B * m.max_len * m.n_channels;
  float* cur = a->x;
  float* nxt = a->x_tmp;
  // one step, directly (src.cur == nullptr) or as part of a capture (per-step parameters through the device cursor)
  cudaStream_t ls = s;  // the stream run_step launches on: the caller's, or the capture stream
  auto run_step = [&](int i, StepSource src) -> int {
    const ffd_step_coef& c = a->coef_host[i];
    FFD_CHECK_ARG(c.dt > 0.f, "step_size must be positive (sde.py:158,239)");
    ffd_forward_args fi = f;
    fi.X = cur;
    fi.time_emb = src.cur ? a->time_emb_table : a->time_emb_table + (int64_t)i * m.d_model;
    const int n_own = a->n_own_host[i];
    if (n_own < 0) {
      fi.mode = FFD_MODE_STOCK;
      fi.n_own = 0;
    } else {
      fi.mode = FFD_MODE_CACHED;
      fi.n_own = n_own;
      fi.own_mask = nullptr;  // the sampler's sets are prefixes (caching.py:164-181)
      FFD_CHECK_ARG(n_own <= m.max_len && fi.k_cache && fi.v_cache, "bad cached step %d", i);
    }
    if (!(a->store_crf_host && a->store_crf_host[i])) fi.crf_out = nullptr;
    int rc2;
    if ((rc2 = forward_hidden(fi, ar, ws, ls, a->event, i, src.cur ? &src.cur->ctr : nullptr))) return rc2;
    const float* z = a->noise ? (src.cur ? a->noise : a->noise + (int64_t)(i - a->noise_step0) * n_elem) : nullptr;
    if (a->fresca && !(a->fresca->low_scale == 1.0f && a->fresca->high_scale_host[i] == 1.0f)) {
      // score -> FreSca filter -> SDE update (sampler.py:79-101); three more launches, no host sync
      const ffd_fresca_args& fr = *a->fresca;
      float* raw = fr.scratch + fresca_scratch_floats(f.B, m.max_len, m.n_channels);
      float* filt = raw + n_elem;
      if ((rc2 = launch_unembed_step(ws.h, f.weights + ar.unemb_w, f.weights + ar.unemb_b, raw, 0, nullptr, nullptr,
                                     nullptr, nullptr, 0, 0, nullptr, f.B * m.max_len, m.max_len, m.n_channels,
                                     m.d_model, ls)))
        return rc2;
      if ((rc2 = launch_fresca(fr.U, fr.V, raw, filt, fr.scratch, f.B, m.max_len, m.n_channels, fr.low_scale,
                               fr.high_scale_host[i], fr.cutoff_ratio, fr.strategy, ls)))
        return rc2;
      if (f.score)
        FFD_CHECK_CUDA(cudaMemcpyAsync(f.score, filt, sizeof(float) * n_elem, cudaMemcpyDeviceToDevice, ls));
      if ((rc2 = launch_sde_step(&c, a->G, cur, filt, z, a->seed, a->stream_base + (uint64_t)i, nxt, f.B, m.max_len,
                                 m.n_channels, ls)))
        return rc2;
    } else if ((rc2 = launch_unembed_step(ws.h, f.weights + ar.unemb_w, f.weights + ar.unemb_b, f.score, 1, &c, a->G,
                                          cur, z, a->seed, a->stream_base + (uint64_t)i, nxt, f.B * m.max_len,
                                          m.max_len, m.n_channels, m.d_model, ls, src))) {
      return rc2;
    }
    float* t = cur;
    cur = nxt;
    nxt = t;
    if (!src.cur && n_own == 0) g_pure_step_ran = true;
    return FFD_OK;
  };
  // a step that a captured pair may stand for: pure cache, nothing step-specific on the host side
  auto graphable = [&](int i) {
    return a->graph && a->graph_scratch && !a->event && !a->fresca && a->n_own_host[i] == 0 &&
           !(a->store_crf_host && a->store_crf_host[i]) && f.k_cache && f.v_cache;
  };
  StepCursor* d_cur = (StepCursor*)a->graph_scratch;
  ffd_step_coef* d_coefs = a->graph_scratch ? (ffd_step_coef*)((char*)a->graph_scratch + graph_coef_offset()) : nullptr;
  bool tables_up = false;
  int i = a->step_begin;
  while (i < a->step_end) {
    int run = 0;
    while (i + run < a->step_end && graphable(i + run)) ++run;
    run &= ~1;  // whole pairs: the state is back in `cur` after every replay
    if (run < 4 || !g_pure_step_ran) {  // too short to pay for a capture (or nothing warmed up yet): direct launches
      if ((rc = run_step(i, StepSource{nullptr, nullptr}))) return rc;
      ++i;
      continue;
    }
    FFD_CHECK_ARG(a->graph_scratch_bytes >= ffd_graph_scratch_bytes(a->step_end), "graph_scratch too small");
    if (!tables_up) {
      FFD_CHECK_CUDA(cudaMemcpyAsync(d_coefs, a->coef_host, sizeof(ffd_step_coef) * a->step_end, cudaMemcpyHostToDevice, s));
      tables_up = true;
    }
    const StepCursor hc = {i, a->noise_step0, a->seed, a->stream_base};
    FFD_CHECK_CUDA(cudaMemcpyAsync(d_cur, &hc, sizeof(hc), cudaMemcpyHostToDevice, s));
    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.dims = m, key.precision = f.precision, key.B = f.B, key.weights = f.weights, key.packed = f.packed;
    key.workspace = f.workspace, key.k_cache = f.k_cache, key.v_cache = f.v_cache, key.score = f.score, key.G = a->G;
    key.te = a->time_emb_table, key.noise = a->noise, key.x = cur, key.x_tmp = nxt, key.scratch = a->graph_scratch;
    key.fused = tc_layer_enabled(m) ? 1 : 0;
    GraphEntry* ge = nullptr;
    for (auto& e : g_graphs)
      if (e.used && memcmp(&e.key, &key, sizeof(key)) == 0) ge = &e;
    if (!ge) {
      // capture two steps (a pure-cache step has run directly before: every lazy initialisation is done)
      // on a stream of the library's own: the caller's may be the legacy default stream, which cannot capture
      static cudaStream_t cap = nullptr;
      if (!cap) FFD_CHECK_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
      const unsigned long long l0 = g_launches;
      FFD_CHECK_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed));
      const StepSource src = {d_cur, d_coefs};
      int rcc = FFD_OK;
      ls = cap;
      for (int k = 0; k < 2 && rcc == FFD_OK; ++k) {
        rcc = run_step(i + k, src);
        if (rcc == FFD_OK) rcc = launch_step_advance(d_cur, cap);
      }
      ls = s;
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
      if (rcc != FFD_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        if (rcc == FFD_OK) set_error("CUDA graph capture of the step loop failed: %s", cudaGetErrorString(ce));
        return rcc != FFD_OK ? rcc : FFD_ERR_CUDA;
      }
      GraphEntry& slot = g_graphs[g_graph_next];
      g_graph_next = (g_graph_next + 1) % 8;
      if (slot.used) cudaGraphExecDestroy(slot.exec);
      slot.used = false;
      FFD_CHECK_CUDA(cudaGraphInstantiate(&slot.exec, graph, 0));
      cudaGraphDestroy(graph);
      slot.key = key;
      slot.launches = g_launches - l0;
      g_launches = l0;  // nothing ran during the capture; replays are counted below
      slot.used = true;
      ge = &slot;
    }
    for (int r = 0; r < run / 2; ++r) {
      FFD_CHECK_CUDA(cudaGraphLaunch(ge->exec, s));
      __atomic_fetch_add(&g_launches, ge->launches, __ATOMIC_RELAXED);
    }
    i += run;
  }
  if (cur != a->x)
    FFD_CHECK_CUDA(cudaMemcpyAsync(a->x, cur, sizeof(float) * n_elem, cudaMemcpyDeviceToDevice, s));
  return FFD_OK;
}

}  // extern "C"
