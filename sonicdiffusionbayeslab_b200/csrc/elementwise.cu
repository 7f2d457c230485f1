// HBM-bound fused elementwise kernels of the sampling loop:
//   * latent_update : classifier-free-guidance combine + any linear multistep scheduler update
//                     (DDIM / DPM-Solver(++) 1st-3rd order (+SDE) / LCM / PLMS) + x0 prediction
//                     + history write, one pass, fp32 math, one rounding per output.
//     Replaces /root/reference/src/models.py:238-242 + :253-255 (scheduler.step) and the
//     10-25 tiny ATen launches per step they cause.
//   * layout helpers (NCHW<->NHWC, nearest 2x upsample, 3x3 im2col for the stride-2 and the 8-channel input convolutions) and the M=1 GEMVs of
//     the timestep-embedding path.
#include "ops.cuh"

namespace sonic {

namespace {

template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  // plain (coherent) load: for tensors an OUTPUT of the same launch may alias (sample / history / noise).  The
  // non-coherent path (ld.global.nc) is only defined for data no thread writes during the kernel.
  static __device__ __forceinline__ void loadc(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  static __device__ __forceinline__ float round(float x) { return x; }
  static __device__ __forceinline__ float ld1(const float* p) { return *p; }
  static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w[j]); f[2 * j + 1] = bf16_hi(w[j]); }
  }
  static __device__ __forceinline__ void loadc(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w[j]); f[2 * j + 1] = bf16_hi(w[j]); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                                               pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
  static __device__ __forceinline__ float round(float x) { return round_bf16(x); }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <typename T>
struct UpdatePtrs {
  const T* eu; const T* et; const T* x; const T* h1; const T* h2; const T* h3; const T* z;
  T* ox; T* om; T* o0;
};

template <typename T>
__device__ __forceinline__ void update_math(const UpdateCoeffs& k, bool cfg, float eu, float et, float x,
                                            float h1, float h2, float h3, float z, float& ox, float& om,
                                            float& o0) {
  // guidance in fp32, then round to the model dtype like the reference's noise_pred tensor
  float e = cfg ? eu + k.guidance * (et - eu) : eu;
  e = Vec8<T>::round(e);
  om = Vec8<T>::round(k.m_x * x + k.m_e * e);        // converted model output kept in history
  o0 = k.x0_x * x + k.x0_e * e;                      // predicted original sample / denoised
  ox = k.c_x * x + k.c_e * e + k.c_m0 * om + k.c_h1 * h1 + k.c_h2 * h2 + k.c_h3 * h3 + k.c_z * z;
}

template <typename T>
__global__ void __launch_bounds__(256)
latent_update_kernel(const UpdateCoeffs k, const UpdatePtrs<T> p, long n, long n_x0) {
  const long nvec = n / 8;
  const bool cfg = p.et != nullptr;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long o = i * 8;
    float eu[8], et[8], x[8], h1[8], h2[8], h3[8], z[8];
    Vec8<T>::load(p.eu + o, eu);
    if (cfg) Vec8<T>::load(p.et + o, et);
    Vec8<T>::loadc(p.x + o, x);                    // may alias out_sample / out_m0: coherent loads
    if (p.h1) Vec8<T>::loadc(p.h1 + o, h1);
    if (p.h2) Vec8<T>::loadc(p.h2 + o, h2);
    if (p.h3) Vec8<T>::loadc(p.h3 + o, h3);
    if (p.z) Vec8<T>::loadc(p.z + o, z);
    float ox[8], om[8], o0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      update_math<T>(k, cfg, eu[j], cfg ? et[j] : 0.f, x[j], p.h1 ? h1[j] : 0.f, p.h2 ? h2[j] : 0.f,
                     p.h3 ? h3[j] : 0.f, p.z ? z[j] : 0.f, ox[j], om[j], o0[j]);
    // all loads of this vector precede the stores, so out_sample may alias sample / history
    if (p.ox) Vec8<T>::store(p.ox + o, ox);
    if (p.om) Vec8<T>::store(p.om + o, om);
    if (p.o0 && o < n_x0) Vec8<T>::store(p.o0 + o, o0);      // n_x0: multiple of 8 or >= n (checked by the launcher)
  }
  // scalar tail (n not a multiple of 8)
  for (long i = nvec * 8 + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float ox, om, o0;
    update_math<T>(k, cfg, Vec8<T>::ld1(p.eu + i), cfg ? Vec8<T>::ld1(p.et + i) : 0.f, Vec8<T>::ld1(p.x + i),
                   p.h1 ? Vec8<T>::ld1(p.h1 + i) : 0.f, p.h2 ? Vec8<T>::ld1(p.h2 + i) : 0.f,
                   p.h3 ? Vec8<T>::ld1(p.h3 + i) : 0.f, p.z ? Vec8<T>::ld1(p.z + i) : 0.f, ox, om, o0);
    if (p.ox) Vec8<T>::st1(p.ox + i, ox);
    if (p.om) Vec8<T>::st1(p.om + i, om);
    if (p.o0 && i < n_x0) Vec8<T>::st1(p.o0 + i, o0);
  }
}

template <typename T>
int launch_update(const UpdateCoeffs& k, const void* eu, const void* et, const void* x, const void* h1,
                  const void* h2, const void* h3, const void* z, void* ox, void* om, void* o0, long n, long n_x0,
                  cudaStream_t stream) {
  UpdatePtrs<T> p{static_cast<const T*>(eu), static_cast<const T*>(et), static_cast<const T*>(x),
                  static_cast<const T*>(h1), static_cast<const T*>(h2), static_cast<const T*>(h3),
                  static_cast<const T*>(z),  static_cast<T*>(ox),       static_cast<T*>(om),
                  static_cast<T*>(o0)};
  const long nvec = std::max<long>(1, n / 8);
  // multiples of the SM count; two resident 256-thread CTAs per SM cover the small latents
  const int blocks = static_cast<int>(std::min<long>((nvec + 255) / 256, 148L * 8));
  latent_update_kernel<T><<<std::max(blocks, 1), 256, 0, stream>>>(k, p, n, n_x0);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------ x0 clipping / dynamic thresholding
// The guided model output and the x0 prediction exactly as update_math forms them, with x0 rounded to the model dtype
// like the reference's x0_pred tensor before it is clipped / thresholded.
template <typename T>
__device__ __forceinline__ float guided_x0(const UpdateCoeffs& k, const T* eu, const T* et, const T* x, long i, float& e,
                                           float& xv) {
  const float u = Vec8<T>::ld1(eu + i);
  e = et ? u + k.guidance * (Vec8<T>::ld1(et + i) - u) : u;
  e = Vec8<T>::round(e);
  xv = Vec8<T>::ld1(x + i);
  return Vec8<T>::round(k.x0_x * xv + k.x0_e * e);
}

template <typename T>
__global__ void __launch_bounds__(256)
latent_update_post_kernel(const UpdateCoeffs k, const UpdatePtrs<T> p, const X0Post q, long n, long n_x0) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float e, x;
    float x0 = guided_x0<T>(k, p.eu, p.et, p.x, i, e, x);
    if (q.mode == 2) {
      const float s = q.thr[i / q.n_per_image];
      x0 = fminf(fmaxf(x0, -s), s) / s;
    } else {
      x0 = fminf(fmaxf(x0, -q.clip), q.clip);
    }
    const float om = Vec8<T>::round(q.p_x * x + q.p_0 * x0);
    const float ox = k.c_x * x + k.c_e * e + k.c_m0 * om + (p.h1 ? k.c_h1 * Vec8<T>::ld1(p.h1 + i) : 0.f) +
                     (p.h2 ? k.c_h2 * Vec8<T>::ld1(p.h2 + i) : 0.f) + (p.h3 ? k.c_h3 * Vec8<T>::ld1(p.h3 + i) : 0.f) +
                     (p.z ? k.c_z * Vec8<T>::ld1(p.z + i) : 0.f);
    // every load of element i precedes its stores, so out_sample may alias sample / history
    if (p.ox) Vec8<T>::st1(p.ox + i, ox);
    if (p.om) Vec8<T>::st1(p.om + i, om);
    if (p.o0 && i < n_x0) Vec8<T>::st1(p.o0 + i, x0);
  }
}

// k-th smallest (0-based) of n non-negative floats held as their bit patterns: most-significant-digit-first radix
// select, 8 bits per pass over a 256-bin shared histogram.  Called by the whole CTA.
__device__ float radix_select(const uint32_t* keys, int n, unsigned rank, unsigned* hist, unsigned* sh) {
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t kk = keys[i];
      if ((kk & mask) == prefix) atomicAdd(&hist[(kk >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned cum = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (rank < cum + hist[b]) break;
        cum += hist[b];
      }
      sh[0] = static_cast<unsigned>(b);
      sh[1] = rank - cum;
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    mask |= 255u << shift;
    rank = sh[1];
    __syncthreads();
  }
  return __uint_as_float(prefix);
}

template <typename T>
__global__ void __launch_bounds__(1024)
x0_threshold_kernel(const UpdateCoeffs k, const T* eu, const T* et, const T* x, int n_per_image, float ratio,
                    float max_value, float* __restrict__ thr) {
  extern __shared__ uint32_t keys[];             // |x0| of this image
  __shared__ unsigned hist[256];
  __shared__ unsigned sh[2];
  const long base = static_cast<long>(blockIdx.x) * n_per_image;
  for (int i = threadIdx.x; i < n_per_image; i += blockDim.x) {
    float e, xv;
    keys[i] = __float_as_uint(fabsf(guided_x0<T>(k, eu, et, x, base + i, e, xv)));
  }
  __syncthreads();
  // torch.quantile(|x0|, ratio, dim=1), interpolation "linear": position ratio * (n - 1) in float32
  const float pos = ratio * static_cast<float>(n_per_image - 1);
  const float lo = floorf(pos);
  const float w = pos - lo;
  const float a = radix_select(keys, n_per_image, static_cast<unsigned>(lo), hist, sh);
  const float b = radix_select(keys, n_per_image, static_cast<unsigned>(ceilf(pos)), hist, sh);
  if (threadIdx.x == 0) {
    const float v = w < 0.5f ? a + w * (b - a) : b - (b - a) * (1.0f - w);      // at::lerp
    thr[blockIdx.x] = fminf(fmaxf(v, 1.0f), max_value);
  }
}

// ------------------------------------------------------------------------------ layout helpers
template <typename T>
__global__ void nchw_to_nhwc8_kernel(const T* __restrict__ x, int n_img, int C, int hw, int dup,
                                     __nv_bfloat16* __restrict__ y) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;   // pixel index
  if (i >= static_cast<long>(n_img) * hw) return;
  const int img = static_cast<int>(i / hw), p = static_cast<int>(i % hw);
  float f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    f[c] = c < C ? Vec8<T>::ld1(x + (static_cast<size_t>(img) * C + c) * hw + p) : 0.f;
  const uint4 u = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                             pack_bf16(f[6], f[7]));
  reinterpret_cast<uint4*>(y)[i] = u;
  if (dup) reinterpret_cast<uint4*>(y)[i + static_cast<long>(n_img) * hw] = u;
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int ld, int n_img, int C, int hw,
                                    T* __restrict__ y) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;   // pixel index
  if (i >= static_cast<long>(n_img) * hw) return;
  const int img = static_cast<int>(i / hw), p = static_cast<int>(i % hw);
  for (int c = 0; c < C; ++c)
    Vec8<T>::st1(y + (static_cast<size_t>(img) * C + c) * hw + p, __bfloat162float(x[i * ld + c]));
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n_img, int H, int W,
                                  int vpp) {
  const long total = static_cast<long>(n_img) * 4 * H * W * vpp;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    long pix = i / vpp;
    const int wo = static_cast<int>(pix % (2 * W)); pix /= 2 * W;
    const int ho = static_cast<int>(pix % (2 * H));
    const int img = static_cast<int>(pix / (2 * H));
    y[i] = __ldg(x + ((static_cast<long>(img) * H + ho / 2) * W + wo / 2) * vpp + v);
  }
}

__global__ void im2col3x3_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n_img, int H, int W,
                                 int vpp, int stride) {
  const int Ho = H / stride, Wo = W / stride;
  const long total = static_cast<long>(n_img) * Ho * Wo * 9 * vpp;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    long r = i / vpp;
    const int tap = static_cast<int>(r % 9); r /= 9;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int img = static_cast<int>(r / Ho);
    const int hi = stride * ho + tap / 3 - 1, wi = stride * wo + tap % 3 - 1;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      u = __ldg(x + ((static_cast<long>(img) * H + hi) * W + wi) * vpp + v);
    y[i] = u;
  }
}

// ------------------------------------------------------------------------------ timestep path
__global__ void timestep_embedding_kernel(const float* __restrict__ t_dev, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half) return;
  const float freq = expf(-logf(10000.0f) * static_cast<float>(j) / static_cast<float>(half));
  const float a = t_dev[0] * freq;
  out[j] = cosf(a);
  out[half + j] = sinf(a);
}

// one warp per output row; jobs are located by binary search over row_start
__global__ void __launch_bounds__(256)
gemv_batched_kernel(const GemvJob* __restrict__ jobs, int n_jobs, int total_rows, const float* __restrict__ x,
                    int K, int silu_in) {
  extern __shared__ float xs[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v = x[k];
    xs[k] = silu_in ? v / (1.0f + __expf(-v)) : v;
  }
  __syncthreads();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total_rows) return;
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].row_start <= row) lo = mid; else hi = mid - 1;
  }
  const GemvJob job = jobs[lo];
  const int n = row - job.row_start;
  const uint4* wr = reinterpret_cast<const uint4*>(job.w + static_cast<size_t>(n) * K);
  float acc = 0.f;
  for (int v = lane; v < K / 8; v += 32) {
    const uint4 u = __ldg(wr + v);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc += bf16_lo(w[j]) * xs[v * 8 + 2 * j] + bf16_hi(w[j]) * xs[v * 8 + 2 * j + 1];
  }
  acc = warp_sum(acc);
  if (lane == 0) job.y[n] = acc + (job.bias ? job.bias[n] : 0.f) + (job.add ? job.add[n] : 0.f);
}

}  // namespace

int latent_update_launch(const UpdateCoeffs& k, const void* eps_uncond, const void* eps_text, const void* sample,
                         const void* h1, const void* h2, const void* h3, const void* noise, void* out_sample,
                         void* out_m0, void* out_x0, long n, long n_x0, int dtype, cudaStream_t stream) {
  SONIC_REQUIRE(eps_uncond && sample, "latent_update: eps and sample are required");
  SONIC_REQUIRE(n > 0, "latent_update: n=%ld", n);
  if (n_x0 < 0 || n_x0 > n) n_x0 = n;
  SONIC_REQUIRE(n_x0 == n || n_x0 % 8 == 0, "latent_update: n_x0=%ld must be a multiple of 8 (or n)", n_x0);
  if (dtype == kF32)
    return launch_update<float>(k, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample, out_m0, out_x0, n,
                                n_x0, stream);
  if (dtype == kBF16)
    return launch_update<__nv_bfloat16>(k, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample, out_m0,
                                        out_x0, n, n_x0, stream);
  SONIC_REQUIRE(false, "latent_update: unknown dtype %d", dtype);
}

namespace {
template <typename T>
int launch_update_post(const UpdateCoeffs& k, const X0Post& q, const void* eu, const void* et, const void* x,
                       const void* h1, const void* h2, const void* h3, const void* z, void* ox, void* om, void* o0, long n,
                       long n_x0, cudaStream_t stream) {
  UpdatePtrs<T> p{static_cast<const T*>(eu), static_cast<const T*>(et), static_cast<const T*>(x),
                  static_cast<const T*>(h1), static_cast<const T*>(h2), static_cast<const T*>(h3),
                  static_cast<const T*>(z),  static_cast<T*>(ox),       static_cast<T*>(om),
                  static_cast<T*>(o0)};
  const int blocks = static_cast<int>(std::min<long>((n + 255) / 256, 148L * 8));
  latent_update_post_kernel<T><<<std::max(blocks, 1), 256, 0, stream>>>(k, p, q, n, n_x0);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_threshold(const UpdateCoeffs& k, const void* eu, const void* et, const void* x, int n_img, int n_per_image,
                     float ratio, float max_value, float* thr, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(n_per_image) * sizeof(uint32_t);
  SONIC_CUDA(cudaFuncSetAttribute(x0_threshold_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  x0_threshold_kernel<T><<<n_img, 1024, smem, stream>>>(k, static_cast<const T*>(eu), static_cast<const T*>(et),
                                                       static_cast<const T*>(x), n_per_image, ratio, max_value, thr);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

int latent_update_post_launch(const UpdateCoeffs& k, const X0Post& post, const void* eps_uncond, const void* eps_text,
                              const void* sample, const void* h1, const void* h2, const void* h3, const void* noise,
                              void* out_sample, void* out_m0, void* out_x0, long n, long n_x0, int dtype,
                              cudaStream_t stream) {
  SONIC_REQUIRE(eps_uncond && sample, "latent_update_post: eps and sample are required");
  SONIC_REQUIRE(n > 0, "latent_update_post: n=%ld", n);
  SONIC_REQUIRE(post.mode == 1 || post.mode == 2, "latent_update_post: mode %d (1 = clip, 2 = dynamic threshold)",
                post.mode);
  if (post.mode == 1) SONIC_REQUIRE(post.clip > 0.f, "latent_update_post: clip range %g must be positive", post.clip);
  if (post.mode == 2)
    SONIC_REQUIRE(post.thr && post.n_per_image > 0 && n % post.n_per_image == 0,
                  "latent_update_post: dynamic thresholding needs thr and n_per_image dividing n (n=%ld, per image %ld)",
                  n, post.n_per_image);
  if (n_x0 < 0 || n_x0 > n) n_x0 = n;
  if (dtype == kF32)
    return launch_update_post<float>(k, post, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample, out_m0,
                                     out_x0, n, n_x0, stream);
  if (dtype == kBF16)
    return launch_update_post<__nv_bfloat16>(k, post, eps_uncond, eps_text, sample, h1, h2, h3, noise, out_sample,
                                             out_m0, out_x0, n, n_x0, stream);
  SONIC_REQUIRE(false, "latent_update_post: unknown dtype %d", dtype);
}

int x0_threshold_launch(const UpdateCoeffs& k, const void* eps_uncond, const void* eps_text, const void* sample,
                        int n_img, long n_per_image, float ratio, float max_value, float* thr, int dtype,
                        cudaStream_t stream) {
  SONIC_REQUIRE(eps_uncond && sample && thr, "x0_threshold: eps, sample and thr are required");
  SONIC_REQUIRE(n_img > 0 && n_per_image > 1, "x0_threshold: n_img=%d n_per_image=%ld", n_img, n_per_image);
  SONIC_REQUIRE(n_per_image * 4 <= 200 * 1024, "x0_threshold: %ld elements per image do not fit shared memory",
                n_per_image);
  SONIC_REQUIRE(ratio >= 0.f && ratio <= 1.f && max_value >= 1.f, "x0_threshold: ratio %g / max_value %g", ratio,
                max_value);
  if (dtype == kF32)
    return launch_threshold<float>(k, eps_uncond, eps_text, sample, n_img, static_cast<int>(n_per_image), ratio,
                                   max_value, thr, stream);
  if (dtype == kBF16)
    return launch_threshold<__nv_bfloat16>(k, eps_uncond, eps_text, sample, n_img, static_cast<int>(n_per_image), ratio,
                                           max_value, thr, stream);
  SONIC_REQUIRE(false, "x0_threshold: unknown dtype %d", dtype);
}

int nchw_to_nhwc8_launch(const void* x, int dtype, int n_img, int C, int hw, int dup, void* y,
                         cudaStream_t stream) {
  SONIC_REQUIRE(C <= 8, "nchw_to_nhwc8: C=%d > 8", C);
  const long total = static_cast<long>(n_img) * hw;
  const int blocks = static_cast<int>((total + 255) / 256);
  if (dtype == kF32)
    nchw_to_nhwc8_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(x), n_img, C, hw, dup,
                                                            static_cast<__nv_bfloat16*>(y));
  else
    nchw_to_nhwc8_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), n_img, C,
                                                                    hw, dup, static_cast<__nv_bfloat16*>(y));
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int nhwc_to_nchw_launch(const void* x, int ld, int n_img, int C, int hw, void* y, int dtype, cudaStream_t stream) {
  const long total = static_cast<long>(n_img) * hw;
  const int blocks = static_cast<int>((total + 255) / 256);
  if (dtype == kF32)
    nhwc_to_nchw_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, n_img, C, hw,
                                                           static_cast<float*>(y));
  else
    nhwc_to_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, n_img,
                                                                   C, hw, static_cast<__nv_bfloat16*>(y));
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

// In-place row softmax of a bf16 score matrix: x[r, :] = softmax(scale * x[r, :]).  One warp per row, the
// row lives in registers (cols <= 8192).  Used by the VAE decoder's single-head d=512 attention, which
// is expressed as two GEMMs around this kernel (the flash kernel covers head dims up to 160).
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(__nv_bfloat16* __restrict__ x, int rows, int cols, long ld, float scale_log2) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * ld);
  const int nvec = cols / 8;
  float f[kVecPerLane][8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 u = xr[v];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[i][2 * j] = bf16_lo(w[j]);
        f[i][2 * j + 1] = bf16_hi(w[j]);
        m = fmaxf(m, fmaxf(f[i][2 * j], f[i][2 * j + 1]));
      }
    }
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[i][j] = exp2f((f[i][j] - m) * scale_log2);
        sum += f[i][j];
      }
    }
  }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec)
      xr[v] = make_uint4(pack_bf16(f[i][0] * inv, f[i][1] * inv), pack_bf16(f[i][2] * inv, f[i][3] * inv),
                         pack_bf16(f[i][4] * inv, f[i][5] * inv), pack_bf16(f[i][6] * inv, f[i][7] * inv));
  }
}

int softmax_rows_launch(void* x, int rows, int cols, long ld, float scale, cudaStream_t stream) {
  SONIC_REQUIRE(cols % 8 == 0 && cols <= 8192 && ld % 8 == 0, "softmax_rows: cols=%d unsupported", cols);
  const int vpl = (cols / 8 + 31) / 32;
  const dim3 grid((rows + 7) / 8);
  auto xb = static_cast<__nv_bfloat16*>(x);
  const float s2 = scale * 1.4426950408889634f;
  if (vpl <= 4) softmax_rows_kernel<4><<<grid, 256, 0, stream>>>(xb, rows, cols, ld, s2);
  else if (vpl <= 16) softmax_rows_kernel<16><<<grid, 256, 0, stream>>>(xb, rows, cols, ld, s2);
  else softmax_rows_kernel<32><<<grid, 256, 0, stream>>>(xb, rows, cols, ld, s2);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int upsample2x_launch(const void* x, void* y, int n_img, int H, int W, int C, cudaStream_t stream) {
  SONIC_REQUIRE(C % 8 == 0, "upsample2x: C=%d", C);
  const long total = static_cast<long>(n_img) * 4 * H * W * (C / 8);
  const int blocks = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
  upsample2x_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), n_img, H, W,
                                                C / 8);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int im2col3x3_launch(const void* x, void* y, int n_img, int H, int W, int C, int stride, cudaStream_t stream) {
  SONIC_REQUIRE(C % 8 == 0 && (stride == 1 || stride == 2) && H % stride == 0 && W % stride == 0,
                "im2col3x3: bad shape / stride");
  const long total = static_cast<long>(n_img) * (H / stride) * (W / stride) * 9 * (C / 8);
  const int blocks = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
  im2col3x3_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), n_img, H, W,
                                               C / 8, stride);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int gemv_batched_launch(const GemvJob* jobs_dev, int n_jobs, int total_rows, const float* x, int K, int silu_in,
                        cudaStream_t stream) {
  SONIC_REQUIRE(K % 8 == 0 && K * sizeof(float) <= 48 * 1024, "gemv: K=%d unsupported", K);
  const int blocks = (total_rows + 7) / 8;
  gemv_batched_kernel<<<blocks, 256, K * sizeof(float), stream>>>(jobs_dev, n_jobs, total_rows, x, K, silu_in);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int timestep_embedding_launch(const float* t_dev, int dim, float* out, cudaStream_t stream) {
  timestep_embedding_kernel<<<(dim / 2 + 127) / 128, 128, 0, stream>>>(t_dev, dim, out);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sonic
