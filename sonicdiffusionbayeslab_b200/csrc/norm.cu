// HBM-bound normalisation kernels over channels-last bf16 activations:
//   GroupNorm(32 groups) [+ SiLU] over the channel-concat of up to two tensors, and LayerNorm.
// Both read each element once per pass with 128-bit accesses and reduce with warp shuffles.
#include "ops.cuh"

namespace sonic {

namespace {

constexpr int kGnThreads = 256;

struct GnSrc {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int c0, c1, ld0, ld1;
};

__device__ __forceinline__ uint4 gn_load(const GnSrc& s, size_t pix, int ch) {
  // ch is a multiple of 8; c0 is a multiple of 8, so a vector never straddles the two sources
  const __nv_bfloat16* p = ch < s.c0 ? s.x0 + pix * s.ld0 + ch : s.x1 + pix * s.ld1 + (ch - s.c0);
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// grid (chunks, n_img); each CTA accumulates per-channel sum / sum-of-squares over its pixel
// chunk, folds channels into groups and adds 2 floats per group to stats[n_img][groups][2].
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(GnSrc s, int hw, int groups, float* __restrict__ stats) {
  extern __shared__ float sm[];                  // [ppp][2][C] per-pixel-lane partials
  const int C = s.c0 + s.c1;
  const int vpp = C / 8;                         // vectors per pixel
  const int ppp = kGnThreads / vpp > 0 ? kGnThreads / vpp : 1;   // pixels per pass
  for (int i = threadIdx.x; i < ppp * 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int img = blockIdx.y;
  const int chunk = (hw + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk;
  const int p_end = min(hw, p_begin + chunk);
  if (vpp <= kGnThreads) {
    const int v = threadIdx.x % vpp;
    const int pl = threadIdx.x / vpp;
    if (pl < ppp) {
      float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int p = p_begin + pl; p < p_end; p += ppp) {
        uint4 u = gn_load(s, static_cast<size_t>(img) * hw + p, v * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
          a[2 * j] += lo; q[2 * j] += lo * lo;
          a[2 * j + 1] += hi; q[2 * j + 1] += hi * hi;
        }
      }
      // every (pixel-lane, vector) slot has exactly one writer: no atomics, fixed reduction order
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sm[(pl * 2) * C + v * 8 + j] = a[j];
        sm[(pl * 2 + 1) * C + v * 8 + j] = q[j];
      }
    }
  } else {  // very wide rows (C > 2048): loop over vectors (ppp == 1, one owner per vector)
    float* s_sum = sm;
    float* s_sq = sm + C;
    for (int p = p_begin; p < p_end; ++p)
      for (int v = threadIdx.x; v < vpp; v += blockDim.x) {
        uint4 u = gn_load(s, static_cast<size_t>(img) * hw + p, v * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
          s_sum[v * 8 + 2 * j] += lo; s_sq[v * 8 + 2 * j] += lo * lo;
          s_sum[v * 8 + 2 * j + 1] += hi; s_sq[v * 8 + 2 * j + 1] += hi * hi;
        }
      }
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int pl = 0; pl < ppp; ++pl)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        a += sm[(pl * 2) * C + c];
        q += sm[(pl * 2 + 1) * C + c];
      }
    // per-CTA partials, reduced in a fixed order by the apply kernel: bit-reproducible, no memset
    float* dst = stats + ((static_cast<size_t>(img) * gridDim.x + blockIdx.x) * groups + g) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}

// grid (chunks, n_img): y = act((x - mean) * rstd * gamma + beta), one 8-channel vector per thread.
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(GnSrc s, int hw, int groups, float eps, const float* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                __nv_bfloat16* __restrict__ y) {
  extern __shared__ float sm[];                  // scale[C], shift[C]
  const int C = s.c0 + s.c1;
  const int cpg = C / groups;
  const int img = blockIdx.y;
  const float inv_n = 1.0f / (static_cast<float>(hw) * cpg);
  float* s_mean = sm + 2 * C;                    // [groups]
  float* s_rstd = s_mean + groups;               // [groups]
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    const float* src = stats + (static_cast<size_t>(img) * gridDim.x * groups + g) * 2;
    float a = 0.f, q = 0.f;
    for (int ch = 0; ch < static_cast<int>(gridDim.x); ++ch) {
      a += src[static_cast<size_t>(ch) * groups * 2];
      q += src[static_cast<size_t>(ch) * groups * 2 + 1];
    }
    const float mean = a * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    s_mean[g] = mean;
    s_rstd[g] = rsqrtf(var + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * gamma[c];
    sm[c] = sc;
    sm[C + c] = beta[c] - s_mean[g] * sc;
  }
  __syncthreads();
  const int vpp = C / 8;
  const int chunk = (hw + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk;
  const int p_end = min(hw, p_begin + chunk);
  const long total = static_cast<long>(p_end - p_begin) * vpp;
  for (long i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = p_begin + static_cast<int>(i / vpp);
    const int v = static_cast<int>(i % vpp);
    const size_t pix = static_cast<size_t>(img) * hw + p;
    uint4 u = gn_load(s, pix, v * 8);
    uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = v * 8 + 2 * j;
      float lo = bf16_lo(w[j]) * sm[c] + sm[C + c];
      float hi = bf16_hi(w[j]) * sm[c + 1] + sm[C + c + 1];
      if (silu) {
        lo = lo / (1.0f + __expf(-lo));
        hi = hi / (1.0f + __expf(-hi));
      }
      w[j] = pack_bf16(lo, hi);
    }
    *reinterpret_cast<uint4*>(y + pix * C + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// One warp per row; the row lives in registers between the statistics and the normalise pass.
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int rows, int C, float eps,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C / 8;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C);
  float f[kVecPerLane][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      uint4 u = __ldg(xr + v);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[i][2 * j] = bf16_lo(w[j]);
        f[i][2 * j + 1] = bf16_hi(w[j]);
        sum += f[i][2 * j] + f[i][2 * j + 1];
      }
    }
  }
  const float mean = warp_sum(sum) / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / C + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * C);
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = pack_bf16((f[i][2 * j] - mean) * rstd * gg[2 * j] + bb[2 * j],
                         (f[i][2 * j + 1] - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1]);
      yr[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

}  // namespace

int groupnorm_launch(const GroupNormOp& op, cudaStream_t stream) {
  const int C = op.c0 + (op.x1 ? op.c1 : 0);
  SONIC_REQUIRE(op.x0 && op.y && op.stats && op.gamma && op.beta, "groupnorm: null operand");
  SONIC_REQUIRE(C % op.groups == 0 && op.c0 % 8 == 0 && C % 8 == 0, "groupnorm: C=%d groups=%d unsupported",
                C, op.groups);
  SONIC_REQUIRE(2 * C * sizeof(float) <= 48 * 1024, "groupnorm: C=%d too wide", C);
  GnSrc s{static_cast<const __nv_bfloat16*>(op.x0), static_cast<const __nv_bfloat16*>(op.x1), op.c0,
          op.x1 ? op.c1 : 0, op.ld0 ? op.ld0 : op.c0, op.ld1 ? op.ld1 : op.c1};
  // enough CTAs to fill the machine, but at least ~32 pixels per CTA
  int chunks = std::max(1, std::min(op.hw / 32, (4 * 148 + op.n_img - 1) / op.n_img));
  chunks = std::min(chunks, kGroupNormMaxChunks);
  dim3 grid(chunks, op.n_img);
  const size_t smem = 2 * C * sizeof(float);
  const int ppp = std::max(1, kGnThreads / (C / 8));
  SONIC_REQUIRE(ppp * smem <= 48 * 1024, "groupnorm: C=%d needs too much shared memory", C);
  gn_stats_kernel<<<grid, kGnThreads, ppp * smem, stream>>>(s, op.hw, op.groups, op.stats);
  gn_apply_kernel<<<grid, kGnThreads, smem + 2 * op.groups * sizeof(float), stream>>>(
      s, op.hw, op.groups, op.eps, op.stats, op.gamma, op.beta, op.silu, static_cast<__nv_bfloat16*>(op.y));
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_launch(const void* x, void* y, int rows, int C, float eps, const float* gamma,
                     const float* beta, cudaStream_t stream) {
  SONIC_REQUIRE(C % 8 == 0 && C <= 2048, "layernorm: C=%d unsupported", C);
  const int nvec = C / 8;
  const int vpl = (nvec + 31) / 32;
  const int warps = 8;
  dim3 grid((rows + warps - 1) / warps);
  auto xb = static_cast<const __nv_bfloat16*>(x);
  auto yb = static_cast<__nv_bfloat16*>(y);
  switch (vpl) {
    case 1: layernorm_kernel<1><<<grid, warps * 32, 0, stream>>>(xb, rows, C, eps, gamma, beta, yb); break;
    case 2: layernorm_kernel<2><<<grid, warps * 32, 0, stream>>>(xb, rows, C, eps, gamma, beta, yb); break;
    case 3: layernorm_kernel<3><<<grid, warps * 32, 0, stream>>>(xb, rows, C, eps, gamma, beta, yb); break;
    case 4: layernorm_kernel<4><<<grid, warps * 32, 0, stream>>>(xb, rows, C, eps, gamma, beta, yb); break;
    case 5: layernorm_kernel<5><<<grid, warps * 32, 0, stream>>>(xb, rows, C, eps, gamma, beta, yb); break;
    default: layernorm_kernel<8><<<grid, warps * 32, 0, stream>>>(xb, rows, C, eps, gamma, beta, yb); break;
  }
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sonic
