// Internal launch wrappers shared between translation units.
#pragma once
#include "ffd_common.cuh"

namespace ffd {

// ffd_simt.cu
// step_ctr != nullptr: te is the whole [n_steps, d] table and the row is picked on the device (te + *step_ctr * d)
int launch_embed(const float* X, const float* We, const float* be, const float* pos, const float* te, float* h,
                 int64_t M, int L, int C, int d, cudaStream_t s, const int32_t* step_ctr = nullptr);
int launch_linear(const float* A, int lda, const float* W, int ldw, const float* bias, float* Y, int ldy,
                  int64_t M, int N, int K, int relu, cudaStream_t s);
int launch_attention(const float* qkv, int ldq, const float* kc, const float* vc, const int32_t* own_mask,
                     int n_prefix_own, float* out, int64_t B, int L, int d, int H, cudaStream_t s);
int launch_add_ln(const float* res, const float* x, const float* w, const float* b, float* y, int64_t M, int d,
                  cudaStream_t s);
int launch_ffn(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, float* out,
               int64_t M, int d, int F, cudaStream_t s);
int launch_scatter_kv(const float* src, int lds, int k_off, int v_off, const int32_t* own_mask, int n_prefix_own,
                      float* kc, float* vc, int L, int H, int hd, cudaStream_t s);
int launch_unembed_step(const float* h, const float* Wu, const float* bu, float* score_out, int do_step,
                        const ffd_step_coef* coef, const float* G, const float* x, const float* z, uint64_t seed,
                        uint64_t stream_id, float* x_out, int64_t M, int L, int C, int d, cudaStream_t s,
                        StepSource src = StepSource{nullptr, nullptr});
// *ctr += 1 (one thread): closes a step of a captured step loop
int launch_step_advance(StepCursor* cur, cudaStream_t s);
int launch_sde_step(const ffd_step_coef* c, const float* G, const float* x, const float* score, const float* z,
                    uint64_t seed, uint64_t stream_id, float* x_out, int64_t B, int L, int C, cudaStream_t s);
int launch_prior(const float* G, const float* z, uint64_t seed, uint64_t stream_id, float scale, float* x_out,
                 int64_t B, int L, int C, cudaStream_t s);
int launch_philox_normal(uint64_t seed, uint64_t stream_id, float* out, int64_t n, cudaStream_t s);
int launch_posemb_renorm(float* table, int L, int d, float max_norm, cudaStream_t s);
int launch_time_embedding(const float* t, int n, const float* W, const float* dense_w, const float* dense_b, int d,
                          float* out, cudaStream_t s);
int launch_dft_matrix(float* out, int L, int inverse, cudaStream_t s);
int launch_dft_apply(const float* T, const float* x, const float* scale, const float* shift, float* y, int64_t B,
                     int L, int C, cudaStream_t s);

int64_t fresca_scratch_floats(int64_t B, int L, int C);
int launch_fresca(const float* U, const float* V, const float* score, float* out, float* scratch, int64_t B, int L,
                  int C, float low, float high, float cutoff_ratio, int strategy, cudaStream_t s);

int launch_event_intensity(const float* crf, const float* prev, int64_t rows, int d, float tau_0, float* out,
                           cudaStream_t s);

// ffd_event.cu (device-decided E2-CRF schedule)
int launch_event_decide(const ffd_event_args& ev, const float* z, const float* x0, int L, int d, int C, int step,
                        int first, cudaStream_t s);
int launch_event_refresh_kv(const float* Wkv, const float* bkv, const float* h0, const float* weight, float* kc,
                            float* vc, int L, int d, int H, cudaStream_t s);

// ffd_metrics.cu
int launch_metric_project(const float* data, int64_t N, int D, const float* dirs, int R, float* out, cudaStream_t s);
int launch_metric_sort_rows(float* x, int R, int64_t N, cudaStream_t s);
int launch_metric_w2_sorted(const float* a, int64_t N, const float* b, int64_t M, int R, float* out, cudaStream_t s);
int launch_spectral_density(const float* xt, float* out, int64_t B, int L, int C, cudaStream_t s);
int launch_metric_row_std(const float* x, int R, int64_t N, float* out, cudaStream_t s);

// ffd_tc.cu (tcgen05 path)
int64_t tc_packed_bytes(const ffd_dims& m);
int tc_pack_weights(const ffd_dims& m, const float* weights, void* packed, cudaStream_t s);
bool tc_supported(const ffd_dims& m);
// out = LN2(x + W2 relu(W1 x + b1) + b2) for layer `layer`, rows [0, M)
int tc_ffn_ln(const ffd_dims& m, const void* packed, const float* weights, int layer, const float* x, float* y,
              int64_t M, cudaStream_t s);

// ffd_tc_attn.cu (tcgen05 pure-cache attention block)
bool tc_attn_supported(const ffd_dims& m);
int64_t tc_attn_image_bytes(const ffd_dims& m);
int64_t tc_kv_image_bytes(const ffd_dims& m);
int64_t tc_ffn_image_bytes(const ffd_dims& m);
const void* tc_attn_image(const ffd_dims& m, const void* packed, int layer);
int tc_pack_attn(const ffd_dims& m, const float* weights, void* img_base, int64_t layer_stride_bytes,
                 int64_t attn_off_bytes, cudaStream_t s);
// bf16 K/V operand images of all layers/heads from the fp32 sample-0 cache [NL, H, L, hd]
int tc_pack_kv(const ffd_dims& m, const float* kc, const float* vc, void* kvimg, cudaStream_t s);
// the same for one layer (kc / vc / kvimg are the bases of all layers)
int tc_pack_kv_layer(const ffd_dims& m, int layer, const float* kc, const float* vc, void* kvimg, cudaStream_t s);
// y = LN1(x + out_proj(softmax(q K_c^T / sqrt(hd)) V_c)) for layer `layer`, rows [0, M); y may alias x
int tc_attn_cached(const ffd_dims& m, const void* attn_img, const void* kvimg, const float* weights, int layer,
                   const float* x, float* y, int64_t M, cudaStream_t s);

// stock attention block (every sample attends to its own keys): K/V rows are read from kv_src[(b L + key) * ld +
// k_off / v_off + ...] (fp32), packed into per-sample operand images (kvimg_own, tc_kv_own_image_bytes) and fed
// to the same kernel with sample-aligned tiles.  y may alias x.  With kv_src == nullptr (allowed when
// tc_attn_own_fused(m)) the K/V projection runs on the tensor cores inside the image build, straight from x.
int64_t tc_kv_own_image_bytes(const ffd_dims& m, int64_t B);
bool tc_attn_own_fused(const ffd_dims& m);
// kvimg_shared != nullptr (mixed branch): only the tokens of the recompute set (prefix [0, n_prefix) or own_mask)
// keep their own K/V rows, every other key row is taken from the shared sample-0 images (tc_pack_kv) of `layer`.
int tc_attn_own(const ffd_dims& m, const void* attn_img, void* kvimg_own, const float* kv_src, int ld, int k_off,
                int v_off, const float* weights, int layer, const float* x, float* y, int64_t M, cudaStream_t s,
                const void* kvimg_shared = nullptr, int n_prefix = 0, const int32_t* own_mask = nullptr);

// per-layer pieces of the packed weight image
const void* tc_ffn_image(const ffd_dims& m, const void* packed, int layer);       // W1 / W2 128-wide slice images
const void* tc_ffn_half_image(const ffd_dims& m, const void* packed, int layer);  // the same as per-CTA halves

// ffd_tc_layer.cu: ONE kernel per pure-cache layer (attention block + FFN block, MUFU and tensor pipe overlapped)
bool tc_layer_supported(const ffd_dims& m);
bool tc_layer_enabled(const ffd_dims& m);   // supported and not switched off (FFD_FUSED_LAYER=0)
void tc_debug_set_fused_layer(int mode);
int64_t tc_kv96_image_bytes(const ffd_dims& m);
// K/V operand images in chunks of 96 keys (the fused kernel's virtual heads), all layers / one layer
int tc_pack_kv96(const ffd_dims& m, const float* kc, const float* vc, void* kvimg96, cudaStream_t s);
int tc_pack_kv96_layer(const ffd_dims& m, int layer, const float* kc, const float* vc, void* kvimg96, cudaStream_t s);
// y = LN2(y1 + FFN(y1)), y1 = LN1(x + cached attention(x)); y may alias x
int tc_layer_cached(const ffd_dims& m, const void* packed, const void* kvimg96, const float* weights, int layer,
                    const float* x, float* y, int64_t M, cudaStream_t s);

}  // namespace ffd
