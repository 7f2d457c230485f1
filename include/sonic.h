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
 *           out has N/2 columns: value * gelu(gate)).
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
} sonic_gemm_args;
int sonic_conv_gemm(const sonic_gemm_args* args, sonic_stream_t stream);
/* Tile width the library would choose for (N, M) -- needed to pack GEGLU weights. */
int sonic_gemm_block_n(int32_t N, int32_t n_img, int32_t H, int32_t W, int32_t epilogue);

#ifdef __cplusplus
}
#endif
#endif /* SONIC_H_ */
