// extern "C" entry points of libsonic (declared in include/sonic.h).
#include "../../include/sonic.h"

#include "common.cuh"
#include "gemm.cuh"
#include "ops.cuh"

using namespace sonic;

extern "C" {

const char* sonic_last_error(void) { return last_error(); }
const char* sonic_version(void) { return "sonic-b200 0.1 sm_100a"; }

int sonic_conv_gemm(const sonic_gemm_args* a, sonic_stream_t stream) {
  SONIC_REQUIRE(a != nullptr, "sonic_conv_gemm: null args");
  GemmOp op;
  op.a0 = a->a0; op.c0 = a->c0; op.ld0 = a->ld0;
  op.a1 = a->a1; op.c1 = a->c1; op.ld1 = a->ld1;
  op.n_img = a->n_img; op.H = a->H; op.W = a->W;
  op.w = a->w; op.N = a->N; op.taps = a->taps;
  op.bias = a->bias; op.row_bias = a->row_bias;
  op.residual = a->residual; op.ld_res = a->ld_res;
  op.out = a->out; op.ld_out = a->ld_out;
  op.epilogue = a->epilogue; op.block_n = a->block_n; op.gn_partial = a->gn_partial;
  op.ln_stats_out = a->ln_stats_out; op.row_scale = a->row_scale;
  op.stride = a->stride == 2 ? 2 : 1; op.upsample = a->upsample;
  GemmPlan plan;
  if (int rc = gemm_plan(op, &plan)) return rc;
  return gemm_launch(plan, static_cast<cudaStream_t>(stream));
}

int sonic_gemm_block_n(int32_t N, int32_t n_img, int32_t H, int32_t W, int32_t epilogue) {
  return gemm_choose_block_n(N, n_img, H, W, epilogue);
}

int sonic_attention(const sonic_attention_args* a, sonic_stream_t stream) {
  SONIC_REQUIRE(a != nullptr, "sonic_attention: null args");
  AttentionOp op;
  op.q = a->q; op.k = a->k; op.v = a->v; op.o = a->o;
  op.ld_q = a->ld_q; op.ld_k = a->ld_k; op.ld_v = a->ld_v; op.ld_o = a->ld_o;
  op.batch = a->batch; op.heads = a->heads; op.seq_q = a->seq_q; op.seq_k = a->seq_k;
  op.head_dim = a->head_dim; op.scale = a->scale; op.causal = a->causal;
  AttentionPlan* plan = nullptr;
  if (int rc = attention_plan(op, &plan)) return rc;
  int rc = attention_launch(plan, static_cast<cudaStream_t>(stream));
  attention_plan_free(plan);
  return rc;
}

int sonic_groupnorm_silu(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img, int32_t hw,
                         int32_t groups, float eps, const float* gamma, const float* beta, int32_t silu,
                         float* stats, void* y, sonic_stream_t stream) {
  GroupNormOp op;
  op.x0 = x0; op.c0 = c0; op.x1 = x1; op.c1 = c1;
  op.n_img = n_img; op.hw = hw; op.groups = groups; op.eps = eps;
  op.gamma = gamma; op.beta = beta; op.silu = silu; op.stats = stats; op.y = y;
  return groupnorm_launch(op, static_cast<cudaStream_t>(stream));
}

int sonic_groupnorm_fused(const void* x0, int32_t c0, const float* part0, const void* x1, int32_t c1,
                          const float* part1, int32_t n_img, int32_t hw, int32_t groups, float eps,
                          const float* gamma, const float* beta, int32_t silu, float* stats, void* y,
                          sonic_stream_t stream) {
  GroupNormOp op;
  op.x0 = x0; op.c0 = c0; op.x1 = x1; op.c1 = c1; op.part0 = part0; op.part1 = part1;
  op.n_img = n_img; op.hw = hw; op.groups = groups; op.eps = eps;
  op.gamma = gamma; op.beta = beta; op.silu = silu; op.stats = stats; op.y = y;
  SONIC_REQUIRE(part0 != nullptr && (x1 == nullptr || part1 != nullptr), "sonic_groupnorm_fused: null partials");
  return groupnorm_launch(op, static_cast<cudaStream_t>(stream));
}

int sonic_layernorm(const void* x, void* y, int32_t rows, int32_t C, float eps, const float* gamma,
                    const float* beta, sonic_stream_t stream) {
  return layernorm_launch(x, y, rows, C, eps, gamma, beta, static_cast<cudaStream_t>(stream));
}

int sonic_ln_side(const float* partials, int32_t parts, int32_t M, int32_t K, float eps, void* side, float* rstd,
                  sonic_stream_t stream) {
  return ln_side_launch(partials, parts, M, K, eps, side, rstd, static_cast<cudaStream_t>(stream));
}

int sonic_latent_update(const sonic_update_coeffs* k, const void* eps_uncond, const void* eps_text,
                        const void* sample, const void* h1, const void* h2, const void* h3, const void* noise,
                        void* out_sample, void* out_m0, void* out_x0, int64_t n, int32_t dtype,
                        sonic_stream_t stream) {
  SONIC_REQUIRE(k != nullptr, "sonic_latent_update: null coefficients");
  UpdateCoeffs c{k->guidance, k->m_x, k->m_e, k->x0_x, k->x0_e, k->c_x, k->c_e, k->c_m0, k->c_h1, k->c_h2,
                 k->c_h3, k->c_z};
  return latent_update_launch(c, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample, out_m0, out_x0,
                              static_cast<long>(n), static_cast<long>(n), dtype, static_cast<cudaStream_t>(stream));
}

int sonic_latent_update_x0n(const sonic_update_coeffs* k, const void* eps_uncond, const void* eps_text,
                            const void* sample, const void* h1, const void* h2, const void* h3, const void* noise,
                            void* out_sample, void* out_m0, void* out_x0, int64_t n, int64_t n_x0, int32_t dtype,
                            sonic_stream_t stream) {
  SONIC_REQUIRE(k != nullptr, "sonic_latent_update_x0n: null coefficients");
  UpdateCoeffs c{k->guidance, k->m_x, k->m_e, k->x0_x, k->x0_e, k->c_x, k->c_e, k->c_m0, k->c_h1, k->c_h2,
                 k->c_h3, k->c_z};
  return latent_update_launch(c, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample, out_m0, out_x0,
                              static_cast<long>(n), static_cast<long>(n_x0), dtype,
                              static_cast<cudaStream_t>(stream));
}

int sonic_latent_update_post(const sonic_update_coeffs* k, const sonic_x0_post* post, const void* eps_uncond,
                             const void* eps_text, const void* sample, const void* h1, const void* h2, const void* h3,
                             const void* noise, void* out_sample, void* out_m0, void* out_x0, int64_t n, int64_t n_x0,
                             int32_t dtype, sonic_stream_t stream) {
  SONIC_REQUIRE(k != nullptr && post != nullptr, "sonic_latent_update_post: null coefficients");
  UpdateCoeffs c{k->guidance, k->m_x, k->m_e, k->x0_x, k->x0_e, k->c_x, k->c_e, k->c_m0, k->c_h1, k->c_h2,
                 k->c_h3, k->c_z};
  X0Post q{post->mode, post->clip, post->p_x, post->p_0, post->thr, static_cast<long>(post->n_per_image)};
  return latent_update_post_launch(c, q, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample, out_m0, out_x0,
                                   static_cast<long>(n), static_cast<long>(n_x0), dtype,
                                   static_cast<cudaStream_t>(stream));
}

int sonic_x0_threshold(const sonic_update_coeffs* k, const void* eps_uncond, const void* eps_text, const void* sample,
                       int32_t n_img, int64_t n_per_image, float ratio, float max_value, float* thr, int32_t dtype,
                       sonic_stream_t stream) {
  SONIC_REQUIRE(k != nullptr, "sonic_x0_threshold: null coefficients");
  UpdateCoeffs c{k->guidance, k->m_x, k->m_e, k->x0_x, k->x0_e, k->c_x, k->c_e, k->c_m0, k->c_h1, k->c_h2,
                 k->c_h3, k->c_z};
  return x0_threshold_launch(c, eps_uncond, eps_text, sample, n_img, static_cast<long>(n_per_image), ratio, max_value,
                             thr, dtype, static_cast<cudaStream_t>(stream));
}

int sonic_nchw_to_nhwc8(const void* x, int32_t dtype, int32_t n_img, int32_t C, int32_t hw, int32_t dup, void* y,
                        sonic_stream_t stream) {
  return nchw_to_nhwc8_launch(x, dtype, n_img, C, hw, dup, y, static_cast<cudaStream_t>(stream));
}
int sonic_nhwc_to_nchw(const void* x, int32_t ld, int32_t n_img, int32_t C, int32_t hw, void* y, int32_t dtype,
                       sonic_stream_t stream) {
  return nhwc_to_nchw_launch(x, ld, n_img, C, hw, y, dtype, static_cast<cudaStream_t>(stream));
}
int sonic_upsample2x(const void* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C,
                     sonic_stream_t stream) {
  return upsample2x_launch(x, y, n_img, H, W, C, static_cast<cudaStream_t>(stream));
}
int sonic_im2col_s2(const void* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C,
                    sonic_stream_t stream) {
  return im2col3x3_launch(x, y, n_img, H, W, C, 2, static_cast<cudaStream_t>(stream));
}

int sonic_im2col3x3(const void* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride,
                    sonic_stream_t stream) {
  return im2col3x3_launch(x, y, n_img, H, W, C, stride, static_cast<cudaStream_t>(stream));
}

int sonic_clip_preprocess(const void* images, int32_t dtype, int32_t n_img, int32_t H, int32_t W, const int32_t* hb,
                          const int32_t* hk, int32_t hks, const int32_t* vb, const int32_t* vk, int32_t vks,
                          int32_t max_rows, int32_t top, int32_t left, int32_t S, const float* mean3,
                          const float* std3, void* out, int32_t out_mode, int32_t patch, sonic_stream_t stream) {
  return clip_preprocess_launch(images, dtype, n_img, H, W, hb, hk, hks, vb, vk, vks, max_rows, top, left, S, mean3,
                                std3, out, out_mode, patch, static_cast<cudaStream_t>(stream));
}

int sonic_softmax_rows(void* x, int32_t rows, int32_t cols, int64_t ld, float scale, sonic_stream_t stream) {
  SONIC_REQUIRE(x != nullptr, "sonic_softmax_rows: null operand");
  return softmax_rows_launch(x, rows, cols, static_cast<long>(ld), scale, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
