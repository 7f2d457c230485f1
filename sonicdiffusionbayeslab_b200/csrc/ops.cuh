// Host-side launch interfaces of the HBM-bound operators (norm.cu, elementwise.cu) and of the
// attention kernel (attention.cu).  All pointers are device pointers borrowed from the caller.
#pragma once
#include "common.cuh"

#include <algorithm>

namespace sonic {

enum DType { kF32 = 0, kBF16 = 1 };

// GroupNorm scratch: per-CTA partial (sum, sumsq) per group, at most this many CTAs per image.
constexpr int kGroupNormMaxChunks = 160;

struct GroupNormOp {
  const void* x0 = nullptr; int c0 = 0, ld0 = 0;   // NHWC bf16, channels [0,c0)
  const void* x1 = nullptr; int c1 = 0, ld1 = 0;   // optional channel-concat source
  const float* part0 = nullptr;                     // optional per-32-row (sum, sumsq) partials written by the
  const float* part1 = nullptr;                     // producing GEMMs: replaces the statistics pass
  int n_img = 1, hw = 1, groups = 32;
  float eps = 1e-5f;
  const float* gamma = nullptr; const float* beta = nullptr;   // [c0+c1]
  int silu = 1;
  float* stats = nullptr;                           // scratch [n_img][kGroupNormMaxChunks][groups][2] fp32
  void* y = nullptr;                                // [n_img*hw][c0+c1] bf16
};
int groupnorm_launch(const GroupNormOp& op, cudaStream_t stream);
int groupnorm_launch_count(const GroupNormOp& op);   // kernels the call above launches for this operator: 1 (cluster form) or 2

int layernorm_launch(const void* x, void* y, int rows, int C, float eps, const float* gamma,
                     const float* beta, cudaStream_t stream);

// Folded LayerNorm (see gemm.cuh): per-row partials [M][parts][2] -> rstd[M] and the first 16 bytes of every
// 128-byte row of the bf16 side tensor [M][64] (the rest of the row must be zero and is never written).
int ln_side_launch(const float* partials, int parts, int M, int K, float eps, void* side, float* rstd,
                   cudaStream_t stream);

// Coefficients of the fused CFG-combine + scheduler update (see include/sonic.h).
struct UpdateCoeffs {
  float guidance;
  float m_x, m_e;
  float x0_x, x0_e;
  float c_x, c_e, c_m0, c_h1, c_h2, c_h3, c_z;
};
// Non-linear post-processing of the x0 prediction inside the fused update (off in every shipped config; the linear
// kernel above stays the hot path): mode 1 clamps x0 to +-clip (diffusers ``clip_sample``), mode 2 is dynamic
// thresholding (src/schedulers.py:58-59,85-90 -> diffusers ``_threshold_sample``): x0 <- clamp(x0, -s, s) / s with the
// per-image s = thr[img] computed by x0_threshold_launch.  The converted model output is then re-derived from the
// processed x0:  m0 = p_x * x + p_0 * x0'.
struct X0Post {
  int mode;
  float clip;
  float p_x, p_0;
  const float* thr;
  long n_per_image;
};
int latent_update_post_launch(const UpdateCoeffs& k, const X0Post& post, const void* eps_uncond, const void* eps_text,
                              const void* sample, const void* h1, const void* h2, const void* h3, const void* noise,
                              void* out_sample, void* out_m0, void* out_x0, long n, long n_x0, int dtype,
                              cudaStream_t stream);
// thr[img] = clamp(quantile_ratio(|x0| over the image's n_per_image elements), 1, max_value); the quantile is exact
// (radix select of the two neighbouring order statistics, linear interpolation like torch.quantile).
int x0_threshold_launch(const UpdateCoeffs& k, const void* eps_uncond, const void* eps_text, const void* sample,
                        int n_img, long n_per_image, float ratio, float max_value, float* thr, int dtype,
                        cudaStream_t stream);
int latent_update_launch(const UpdateCoeffs& k, const void* eps_uncond, const void* eps_text,
                         const void* sample, const void* h1, const void* h2, const void* h3,
                         const void* noise, void* out_sample, void* out_m0, void* out_x0, long n,
                         long n_x0, int dtype, cudaStream_t stream);   // n_x0: leading elements of x0 to write

// NCHW (fp32 or bf16) latents -> NHWC bf16 padded to 8 channels; `dup` writes each image twice
// (image i and image i + n_img) for the classifier-free-guidance batch.
int nchw_to_nhwc8_launch(const void* x, int dtype, int n_img, int C, int hw, int dup, void* y,
                         cudaStream_t stream);
// [M][ld] bf16 rows (first C channels) -> NCHW tensor of `dtype`.
int nhwc_to_nchw_launch(const void* x, int ld, int n_img, int C, int hw, void* y, int dtype,
                        cudaStream_t stream);
int upsample2x_launch(const void* x, void* y, int n_img, int H, int W, int C, cudaStream_t stream);
// in place: x[r, :cols] = softmax(scale * x[r, :cols]) over bf16 rows of pitch ld
int softmax_rows_launch(void* x, int rows, int cols, long ld, float scale, cudaStream_t stream);
// stride-2, pad-1 3x3 patches: y[n][ho][wo][tap*C + c]  (Ho = H/2, Wo = W/2)
int im2col3x3_launch(const void* x, void* y, int n_img, int H, int W, int C, int stride, cudaStream_t stream);

// Batched GEMV for the timestep path: for every job j, y_j[n] = b_j[n] + add_j[n] + sum_k W_j[n][k]*act(x[k]).
struct GemvJob {
  const __nv_bfloat16* w;     // [N][K]
  const float* bias;          // [N] or null
  const float* add;           // [N] or null (e.g. conv1.bias folded into the time-embedding bias)
  float* y;                   // [N]
  int N;
  int row_start;              // prefix sum of N over jobs (filled by the launcher's caller)
};
int gemv_batched_launch(const GemvJob* jobs_dev, int n_jobs, int total_rows, const float* x, int K,
                        int silu_in, cudaStream_t stream);
// t -> [cos(t f_j) | sin(t f_j)] (flip_sin_to_cos, freq_shift 0), fp32 out[dim]
int timestep_embedding_launch(const float* t_dev, int dim, float* out, cudaStream_t stream);

// CLIP image preprocessing (preprocess.cu): uint8 / float[0,1] images -> PIL-exact bicubic resize + centre crop +
// normalise.  dtype: 0 fp32, 1 bf16 (both quantised with x*255 -> uint8 first), 2 uint8.  mean3 / std3: HOST arrays.
int clip_preprocess_launch(const void* images, int dtype, int n_img, int H, int W, const int* hb, const int* hk, int hks,
                           const int* vb, const int* vk, int vks, int max_rows, int top, int left, int S,
                           const float* mean3, const float* std3, void* out, int out_mode, int patch,
                           cudaStream_t stream);

struct AttentionOp {
  const void* q; const void* k; const void* v;   // bf16; element (b, s, h, d) at ((b*S + s)*ld + h*D + d)
  int ld_q = 0, ld_k = 0, ld_v = 0;
  void* o; int ld_o = 0;
  int batch = 1, heads = 8, seq_q = 0, seq_k = 0, head_dim = 0;
  float scale = 1.f;
  int causal = 0;
};
struct AttentionPlan;
int attention_plan(const AttentionOp& op, AttentionPlan** out);
int attention_launch(const AttentionPlan* plan, cudaStream_t stream);
void attention_plan_free(AttentionPlan* plan);
double attention_flops(const AttentionOp& op);
double attention_plan_flops(const AttentionPlan* plan);

}  // namespace sonic
