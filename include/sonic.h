/* libsonic -- C ABI of the B200-native Stable Diffusion sampling engine.
 *
 * This is the drop-in boundary for the hot path of Kotstantinovskiy/SonicDiffusionBayesLab:
 * the per-timestep body of the denoising loop (reference: src/models.py:211-261) --
 * UNet forward (src/models.py:227-235), classifier-free-guidance combine (:238-242) and the
 * scheduler update (:253-255 -> src/schedulers.py:98-187 and the diffusers DDIM / LCM / PNDM
 * steps).  The reference has no native code, so there is no existing FFI to mirror; the
 * Python plugin objects behind src/registry.py bind these entry points with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; sonic_last_error()
 *     returns a thread-local message for the last failure;
 *   - all pointers are DEVICE pointers borrowed from the caller (torch tensors:
 *     tensor.data_ptr()) unless a parameter is documented as host memory; nothing is freed
 *     by the library except handles it created;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the host, so calls are CUDA-graph capturable;
 *   - activations are bf16, channels-last (NHWC); "ld" arguments are row pitches in elements.
 */
#ifndef SONIC_H_
#define SONIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sonic_stream_t;

const char* sonic_last_error(void);
/* Library / build identification: "sonic-b200 <version> sm_100a". */
const char* sonic_version(void);

/* ---------------------------------------------------------------------------------------
 * Implicit-GEMM operator on tcgen05 tensor cores: 3x3 (pad 1, stride 1) / 1x1 convolution and
 * Linear.  Replaces the cuDNN / cuBLASLt calls reached from UNet2DConditionModel.forward
 * (reference call site src/models.py:227-235).
 *   out[m, n] = epi( sum_{tap,k} A[pixel(m)+offset(tap), k] * w[tap][n][k] + bias[n]
 *                    + row_bias[image(m)][n] ) + residual[m, n]
 * A = channel-concat of (a0[..., :c0], a1[..., :c1]) (a1 may be NULL), NHWC bf16.
 * A Linear over [M, K] is expressed as n_img=1, H=1, W=M, taps=1.
 * epilogue: 0 = none; 1 = GEGLU (w rows packed per block_n tile as [value half | gate half],
 *           out has N/2 columns: value * gelu(gate)); 2 = QuickGELU x * sigmoid(1.702 x) (CLIP MLPs).
 */
typedef struct sonic_gemm_args {
  const void* a0; int32_t c0, ld0;
  const void* a1; int32_t c1, ld1;
  int32_t n_img, H, W;
  const void* w; int32_t N, taps;
  const float* bias;
  const float* row_bias;
  const void* residual; int32_t ld_res;
  void* out; int32_t ld_out;
  int32_t epilogue;
  int32_t block_n;       /* 0 = let the library choose */
  /* Optional GroupNorm pre-reduction of the OUTPUT: [ceil(M/32)][n_out][2] fp32 (sum, sum of squares) of the
   * bf16-rounded output over each block of 32 consecutive rows, per channel.  sonic_groupnorm_fused turns
   * these into group statistics, so the separate statistics pass over the tensor disappears. */
  float* gn_partial;
  /* LayerNorm folded into the GEMMs around it -- the 48 LayerNorms of the UNet's transformer blocks
   * (BasicTransformerBlock.norm1/2/3, reached from src/models.py:227-235) make no pass over the activations:
   *   LN(x) W^T + b = rstd_row * (x (gamma .* W)^T - mean_row * s + std_row * b'),  s_n = sum_k gamma_k W_nk, b' = W beta + b.
   * PRODUCER of x: ln_stats_out = [M][2 * ceil(N / block_n)][2] fp32 per-row (sum, sumsq) partials of its
   *   bf16-rounded output, one slot per (n-tile, column half): fixed-order, no atomics.  Pass block_n explicitly.
   * sonic_ln_side() turns the partials into rstd[M] and a bf16 side tensor [M][64] holding (-mean, std) as hi/lo pairs.
   * CONSUMER: a0 = x, a1 = the side tensor (c1 = 64), w = [N][K + 64] = (bf16(gamma .* W) | s, b' as hi/lo pairs, zeros)
   *   (kernels.py fold_layernorm), bias = NULL, row_scale = rstd: the two rank-1 terms ride through the tensor core
   *   as one extra K chunk and the epilogue only multiplies each row by rstd. */
  float* ln_stats_out;
  const float* row_scale;
  /* Strided / upsampled 3x3 convolutions without a materialised im2col or upsampled tensor (taps = 9, one source,
   * no residual / row_bias / GEGLU):
   *   stride = 2   : Downsample2D (stride 2, pad 1).  H, W are the OUTPUT extents, a0 is the 2H x 2W input; the nine
   *                  taps are TMA boxes of four parity views of the input.  w: [9][N][K] as for stride 1.
   *   upsample = 1 : Upsample2D (nearest 2x, then 3x3).  H, W are the SOURCE extents, out is the 2H x 2W result.
   *                  Each output phase (y & 1, x & 1) is a 2x2 convolution of the source with taps summed from the
   *                  3x3 kernel (4/9 of the multiply-adds): w = [16][N][K], phase-major (sonic kernels.py
   *                  pack_upsample_conv_weight); the epilogue stores through four strided views of the output. */
  int32_t stride;        /* 0 / 1 = unit stride */
  int32_t upsample;
} sonic_gemm_args;
int sonic_conv_gemm(const sonic_gemm_args* args, sonic_stream_t stream);
/* Tile width the library would choose for (N, M) -- needed to pack GEGLU weights. */
int sonic_gemm_block_n(int32_t N, int32_t n_img, int32_t H, int32_t W, int32_t epilogue);

/* ---------------------------------------------------------------------------------------
 * Flash attention on tcgen05 (self- and cross-attention of BasicTransformerBlock; replaces
 * F.scaled_dot_product_attention reached from src/models.py:227-235).
 * Element (b, s, h, d) of q/k/v/o lives at ((b*seq + s)*ld + h*head_dim + d); bf16.
 */
typedef struct sonic_attention_args {
  const void* q; const void* k; const void* v; void* o;
  int32_t ld_q, ld_k, ld_v, ld_o;
  int32_t batch, heads, seq_q, seq_k, head_dim;
  float scale;
  int32_t causal;        /* 1: key j of query i is masked when j > i (CLIP text towers); 0: full attention */
} sonic_attention_args;
int sonic_attention(const sonic_attention_args* args, sonic_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * HBM-bound fused kernels.
 * GroupNorm(groups) [+SiLU] over NHWC bf16, optionally over the channel concat (x0 | x1).
 * stats: ZERO-INITIALISED fp32 scratch of SONIC_GROUPNORM_SCRATCH_FLOATS(n_img, groups) floats
 * (per-CTA partial sums folded in a fixed order by the last CTA of each image: results are
 * bit-reproducible; the kernel leaves the scratch ready for the next launch on the same stream).
 * LayerNorm over rows of C bf16.
 */
#define SONIC_GROUPNORM_MAX_CHUNKS 160
#define SONIC_GROUPNORM_SCRATCH_FLOATS(n_img, groups) \
  ((n_img) * ((SONIC_GROUPNORM_MAX_CHUNKS + 1) * (groups) * 2 + 1))
int sonic_groupnorm_silu(const void* x0, int32_t c0, const void* x1, int32_t c1, int32_t n_img,
                         int32_t hw, int32_t groups, float eps, const float* gamma,
                         const float* beta, int32_t silu, float* stats, void* y,
                         sonic_stream_t stream);
/* GroupNorm whose statistics come from the producing GEMMs' `gn_partial` buffers (hw must be a multiple of 32):
 * ONE launch -- a thread-block cluster of up to 8 CTAs per image folds the per-32-row partials in a fixed order
 * (cluster barrier + distributed shared memory: bit-reproducible, no atomics, no global scratch) and normalises its
 * pixel chunks; large tensors of small batches use a finalize kernel followed by the apply kernel above instead.
 * part0 / part1 have row pitch c0 / c1 channels; `stats` is only touched by the two-kernel form. */
int sonic_groupnorm_fused(const void* x0, int32_t c0, const float* part0, const void* x1, int32_t c1,
                          const float* part1, int32_t n_img, int32_t hw, int32_t groups, float eps,
                          const float* gamma, const float* beta, int32_t silu, float* stats, void* y,
                          sonic_stream_t stream);
int sonic_layernorm(const void* x, void* y, int32_t rows, int32_t C, float eps, const float* gamma,
                    const float* beta, sonic_stream_t stream);
/* Folded LayerNorm, the step between producer and consumer GEMM (see sonic_gemm_args): partials [M][parts][2] ->
 * rstd[M] (fp32) and bytes 0..15 of every 128-byte row of `side` ([M][64] bf16, ZERO-INITIALISED by the caller:
 * columns 8..63 are never written).  K = LayerNorm width. */
int sonic_ln_side(const float* partials, int32_t parts, int32_t M, int32_t K, float eps, void* side, float* rstd,
                  sonic_stream_t stream);

/* Fused classifier-free-guidance combine + scheduler update + x0 prediction + history write:
 * one kernel for the whole of src/models.py:238-242 and :253-255 (scheduler.step of
 * src/schedulers.py:98-187 and of the diffusers DDIM / LCM / PNDM schedulers).  Every
 * scheduler on the path is a linear multistep rule, so the host reduces it to coefficients:
 *   e   = eps_text ? eps_uncond + guidance*(eps_text - eps_uncond) : eps_uncond  (rounded to dtype)
 *   m0  = m_x*x + m_e*e            (converted model output; written to out_m0, rounded to dtype)
 *   x0  = x0_x*x + x0_e*e          (pred_original_sample / denoised; written to out_x0)
 *   x'  = c_x*x + c_e*e + c_m0*m0 + c_h1*h1 + c_h2*h2 + c_h3*h3 + c_z*z   (written to out_sample)
 * Pointers whose coefficient is zero may be NULL; out_* may be NULL; out_sample may alias
 * sample or a history tensor.  dtype: 0 = fp32, 1 = bf16 (all tensors).  n = element count.
 */
typedef struct sonic_update_coeffs {
  float guidance;
  float m_x, m_e;
  float x0_x, x0_e;
  float c_x, c_e, c_m0, c_h1, c_h2, c_h3, c_z;
} sonic_update_coeffs;
int sonic_latent_update(const sonic_update_coeffs* k, const void* eps_uncond, const void* eps_text,
                        const void* sample, const void* h1, const void* h2, const void* h3,
                        const void* noise, void* out_sample, void* out_m0, void* out_x0, int64_t n,
                        int32_t dtype, sonic_stream_t stream);
/* Same launch, but only the first n_x0 elements of x0 are written (n_x0 a multiple of 8): the reference keeps
 * x0_pred[0] only (src/models.py:257-261 `x0_preds.append(x0_pred[0].unsqueeze(0))`), so the pipelines pass one
 * image's worth and save the other B-1 writes. */
int sonic_latent_update_x0n(const sonic_update_coeffs* k, const void* eps_uncond, const void* eps_text,
                            const void* sample, const void* h1, const void* h2, const void* h3,
                            const void* noise, void* out_sample, void* out_m0, void* out_x0, int64_t n,
                            int64_t n_x0, int32_t dtype, sonic_stream_t stream);

/* The same fused update with a NON-LINEAR post-processing of the x0 prediction -- the branches of
 * /root/reference/src/schedulers.py:58-59 and :85-90 (`thresholding`, diffusers `_threshold_sample`) and diffusers'
 * `clip_sample` (DDIM step, item 4); off in every shipped config, so this is a separate kernel and the linear one above
 * stays the hot path:
 *   x0  = round_dtype(x0_x*x + x0_e*e)
 *   mode 1:  x0' = clamp(x0, -clip, clip)
 *   mode 2:  x0' = clamp(x0, -s, s) / s,  s = thr[image]  (sonic_x0_threshold below; n_per_image elements per image)
 *   m0  = p_x*x + p_0*x0'   (the converted model output re-derived from the processed x0: p = (0, 1) for the `++`
 *                            algorithms and DDIM, (1/sigma_t, -alpha_t/sigma_t) for `dpmsolver` / `sde-dpmsolver`)
 *   x'  as in sonic_latent_update; out_x0 receives x0'. */
typedef struct sonic_x0_post {
  int32_t mode;
  float clip;
  float p_x, p_0;
  const float* thr;
  int64_t n_per_image;
} sonic_x0_post;
int sonic_latent_update_post(const sonic_update_coeffs* k, const sonic_x0_post* post, const void* eps_uncond,
                             const void* eps_text, const void* sample, const void* h1, const void* h2, const void* h3,
                             const void* noise, void* out_sample, void* out_m0, void* out_x0, int64_t n, int64_t n_x0,
                             int32_t dtype, sonic_stream_t stream);
/* Dynamic-thresholding scale per image: thr[i] = clamp(quantile(|x0_i|, ratio), 1, max_value) with x0 as above and the
 * quantile of torch.quantile (exact order statistics, linear interpolation at ratio*(n-1) in float32).  One CTA per
 * image; n_per_image*4 bytes of shared memory (<= 200 KiB). */
int sonic_x0_threshold(const sonic_update_coeffs* k, const void* eps_uncond, const void* eps_text, const void* sample,
                       int32_t n_img, int64_t n_per_image, float ratio, float max_value, float* thr, int32_t dtype,
                       sonic_stream_t stream);

/* Layout helpers used around the UNet (exposed for tests). */
int sonic_nchw_to_nhwc8(const void* x, int32_t dtype, int32_t n_img, int32_t C, int32_t hw,
                        int32_t dup, void* y, sonic_stream_t stream);
int sonic_nhwc_to_nchw(const void* x, int32_t ld, int32_t n_img, int32_t C, int32_t hw, void* y,
                       int32_t dtype, sonic_stream_t stream);
int sonic_upsample2x(const void* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C,
                     sonic_stream_t stream);
int sonic_im2col_s2(const void* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C,
                    sonic_stream_t stream);
/* y[n_img*(H/stride)*(W/stride), 9*C] = 3x3 patches (tap-major, zero padding 1) of NHWC x; stride 1 or 2.  Stride 1
 * feeds conv_in (4 -> 8 padded input channels: nine 16-byte TMA rows per pixel would be row-rate-bound, one
 * 144-byte row is not); reference call site src/models.py:227-235 -> UNet2DConditionModel.conv_in. */
int sonic_im2col3x3(const void* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t stride,
                    sonic_stream_t stream);
/* In-place row softmax of bf16 scores: x[r, :cols] = softmax(scale * x[r, :cols]); rows of pitch ld elements.
 * With two sonic_conv_gemm calls this is the single-head d=512 attention of the VAE decoder
 * (reference call site src/models.py:288-302 -> AutoencoderKL.decode). */
int sonic_softmax_rows(void* x, int32_t rows, int32_t cols, int64_t ld, float scale, sonic_stream_t stream);

/* CLIP image preprocessing of the CLIP-score metric (reference: src/metrics/metrics.py:25-41 -> torchmetrics
 * CLIPScore -> HF CLIPImageProcessor on PIL, the host-side hot spot of SURVEY.md row a12; images produced at
 * src/experiments/base_experiment.py:198-201).  ONE kernel: [quantise x*255 -> uint8] -> PIL's two-pass antialiased
 * bicubic resize (uint8 intermediate, 22-bit fixed-point coefficients: bit-identical to Pillow) -> centre crop ->
 * * 1/255 -> (x - mean) / std.
 *   images   [n_img][3][H][W], dtype 0 = fp32 in [0,1], 1 = bf16 in [0,1], 2 = uint8
 *   hb / hk  horizontal bounds [nw][2] (first input column, count) and coefficients [nw][hks] of the resized width
 *   vb / vk  the same for the resized height; max_rows >= input rows any 16-output-row tile needs
 *   top/left centre-crop offset in the resized image; S = output size (224)
 *   mean3 / std3  HOST pointers to 3 floats
 *   out      out_mode 0: fp32 [n_img][3][S][S];  out_mode 1: bf16 [n_img*(S/patch)^2][3*patch*patch], the A operand
 *            of the ViT patch-embedding GEMM (row = image, patch row, patch column; column = channel, y, x). */
int sonic_clip_preprocess(const void* images, int32_t dtype, int32_t n_img, int32_t H, int32_t W, const int32_t* hb,
                          const int32_t* hk, int32_t hks, const int32_t* vb, const int32_t* vk, int32_t vks,
                          int32_t max_rows, int32_t top, int32_t left, int32_t S, const float* mean3,
                          const float* std3, void* out, int32_t out_mode, int32_t patch, sonic_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Launch plans: the native runtime under the engine.  The host records every operator of a
 * network once (TMA descriptors and tile shapes are resolved at record time); running the plan
 * is a native launch loop on `stream`, or -- after sonic_plan_capture -- one CUDA-graph replay.
 * A plan is what replaces one `self.unet(...)` call (src/models.py:227-235); the DeepCache
 * cached step (src/experiments/deep_cache.py:24-29) is a second, shorter plan over the same
 * buffers.  The `add` functions take the same arguments as the eager entry points above.
 */
typedef void* sonic_plan_t;
int sonic_plan_create(sonic_plan_t* out);
int sonic_plan_destroy(sonic_plan_t plan);
int sonic_plan_add_conv_gemm(sonic_plan_t plan, const sonic_gemm_args* args);
int sonic_plan_add_attention(sonic_plan_t plan, const sonic_attention_args* args);
int sonic_plan_add_groupnorm(sonic_plan_t plan, const void* x0, int32_t c0, const void* x1, int32_t c1,
                             int32_t n_img, int32_t hw, int32_t groups, float eps, const float* gamma,
                             const float* beta, int32_t silu, float* stats, void* y);
int sonic_plan_add_groupnorm_fused(sonic_plan_t plan, const void* x0, int32_t c0, const float* part0,
                                   const void* x1, int32_t c1, const float* part1, int32_t n_img, int32_t hw,
                                   int32_t groups, float eps, const float* gamma, const float* beta,
                                   int32_t silu, float* stats, void* y);
int sonic_plan_add_layernorm(sonic_plan_t plan, const void* x, void* y, int32_t rows, int32_t C,
                             float eps, const float* gamma, const float* beta);
int sonic_plan_add_nchw_to_nhwc8(sonic_plan_t plan, const void* x, int32_t dtype, int32_t n_img,
                                 int32_t C, int32_t hw, int32_t dup, void* y);
int sonic_plan_add_nhwc_to_nchw(sonic_plan_t plan, const void* x, int32_t ld, int32_t n_img, int32_t C,
                                int32_t hw, void* y, int32_t dtype);
int sonic_plan_add_upsample2x(sonic_plan_t plan, const void* x, void* y, int32_t n_img, int32_t H,
                              int32_t W, int32_t C);
int sonic_plan_add_im2col_s2(sonic_plan_t plan, const void* x, void* y, int32_t n_img, int32_t H,
                             int32_t W, int32_t C);
int sonic_plan_add_im2col3x3(sonic_plan_t plan, const void* x, void* y, int32_t n_img, int32_t H,
                             int32_t W, int32_t C, int32_t stride);
int sonic_plan_add_ln_side(sonic_plan_t plan, const float* partials, int32_t parts, int32_t M, int32_t K, float eps,
                           void* side, float* rstd);
/* sonic_softmax_rows as a plan operator (the VAE decoder's single-head attention, src/models.py:288-302). */
int sonic_plan_add_softmax_rows(sonic_plan_t plan, void* x, int32_t rows, int32_t cols, int64_t ld, float scale);
/* out[dim] = [cos(t f_j) | sin(t f_j)], t read from DEVICE memory at run time (graph-safe). */
int sonic_plan_add_timestep_embedding(sonic_plan_t plan, const float* t_dev, int32_t dim, float* out);
/* Batched M=1 GEMV: y_j = bias_j + add_j + W_j[N_j x K] * act(x), act = SiLU if silu_in.
 * The pointer tables are HOST arrays of n_jobs device pointers (bias/add tables may be NULL). */
int sonic_plan_add_gemv(sonic_plan_t plan, int32_t n_jobs, const void* const* w,
                        const float* const* bias, const float* const* add, float* const* y,
                        const int32_t* N, const float* x, int32_t K, int32_t silu_in);
int sonic_plan_run(sonic_plan_t plan, sonic_stream_t stream);
/* Capture the plan into a CUDA graph on `stream` (non-default); later runs replay it. */
int sonic_plan_capture(sonic_plan_t plan, sonic_stream_t stream);
/* Kernel launches per run and algorithmic FLOPs (2*M*N*K of every GEMM/conv/attention). */
int sonic_plan_stats(sonic_plan_t plan, int32_t* n_launches, double* flops);
/* One eager run with a CUDA event between operators: ms[i] = device time of operator i,
 * kinds[i]: 0 gemm/conv, 1 attention, 2 groupnorm, 3 layernorm, 4 layout, 5 timestep path;
 * flops[i] = algorithmic FLOPs (0 for HBM-bound operators).  Used by bench.py for the roofline. */
int sonic_plan_profile(sonic_plan_t plan, sonic_stream_t stream, int32_t max_ops, float* ms,
                       int32_t* kinds, double* flops, int32_t* n_ops);

#ifdef __cplusplus
}
#endif
#endif /* SONIC_H_ */
