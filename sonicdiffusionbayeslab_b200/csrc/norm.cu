// HBM-bound normalisation kernels over channels-last bf16 activations:
//   GroupNorm(32 groups) [+ SiLU] over the channel-concat of up to two tensors, and LayerNorm.
// Both read each element once per pass with 128-bit accesses; GroupNorm statistics are reduced in
// a fixed order, so results are bit-reproducible without atomics on the data path.  Three forms:
//   gn_cluster_kernel                  statistics from the producing GEMMs' epilogue partials, ONE launch (a cluster of
//                                      up to 8 CTAs per image): every GroupNorm of the UNet at the bench batch
//   gn_finalize_kernel + gn_apply      the same partials, two launches (large tensors of small batches: the VAE)
//   gn_stats_kernel + gn_apply         a statistics pass over the tensor (no producer partials)
#include "ops.cuh"

#include <algorithm>
#include <cstdlib>

namespace sonic {

namespace {

constexpr int kGnThreads = 512;

struct GnSrc {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int c0, c1, ld0, ld1;
};

__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ uint4 gn_load(const GnSrc& s, size_t pix, int ch) {
  // ch is a multiple of 8; c0 is a multiple of 8, so a vector never straddles the two sources
  const __nv_bfloat16* p = ch < s.c0 ? s.x0 + pix * s.ld0 + ch : s.x1 + pix * s.ld1 + (ch - s.c0);
  return __ldg(reinterpret_cast<const uint4*>(p));
}
// The cluster kernel can be launched while the kernel that wrote x is still draining (SONIC_PDL): the non-coherent
// path (ld.global.nc) is only defined for data nobody writes during the reader's lifetime, so it reads through L2.
__device__ __forceinline__ uint4 gn_load_cg(const GnSrc& s, size_t pix, int ch) {
  const __nv_bfloat16* p = ch < s.c0 ? s.x0 + pix * s.ld0 + ch : s.x1 + pix * s.ld1 + (ch - s.c0);
  return __ldcg(reinterpret_cast<const uint4*>(p));
}

// Scratch layout (floats): partial[n_img][kGroupNormMaxChunks][groups][2] | final[n_img][groups][2]
//                          | ticket[n_img] (int, zero between launches)
__device__ __forceinline__ float* gn_final(float* stats, int n_img, int groups) {
  return stats + static_cast<size_t>(n_img) * kGroupNormMaxChunks * groups * 2;
}

// grid (chunks, n_img).  Thread (pl, v): 8-channel vector v of every ppp-th pixel of the chunk.
__global__ void __launch_bounds__(kGnThreads, 2)
gn_stats_kernel(GnSrc s, int hw, int groups, float eps, float* __restrict__ stats) {
  extern __shared__ float sm[];                  // [ppp][2][C] per-pixel-lane partials
  __shared__ int s_last;
  const int C = s.c0 + s.c1;
  const int vpp = C / 8;                         // vectors per pixel
  const int ppp = kGnThreads / vpp;              // pixel lanes
  const int img = blockIdx.y, n_img = gridDim.y;
  const int chunk = (hw + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk;
  const int p_end = min(hw, p_begin + chunk);
  const int v = threadIdx.x % vpp;
  const int pl = threadIdx.x / vpp;
  if (pl < ppp) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t base = static_cast<size_t>(img) * hw;
    auto acc = [&](const uint4& u) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
        a[2 * j] += lo; q[2 * j] = fmaf(lo, lo, q[2 * j]);
        a[2 * j + 1] += hi; q[2 * j + 1] = fmaf(hi, hi, q[2 * j + 1]);
      }
    };
    // eight 16-byte loads in flight per thread; the tail is predicated inside the same batch (a separate
    // one-pixel-at-a-time tail loop serialised on memory latency and dominated the small feature maps)
    for (int p = p_begin + pl; p < p_end; p += 8 * ppp) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        u[k] = p + k * ppp < p_end ? gn_load(s, base + p + k * ppp, v * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc(u[k]);               // zeros add nothing to sum / sum of squares
    }
    // every (pixel-lane, vector) slot has exactly one writer: no atomics, fixed reduction order
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[(pl * 2) * C + v * 8 + j] = a[j];
      sm[(pl * 2 + 1) * C + v * 8 + j] = q[j];
    }
  }
  __syncthreads();
  // Fold the per-pixel-lane partials: 2 * groups outputs, 8 threads each (a fixed assignment and a fixed
  // shuffle tree, so the result is still bit-reproducible).  The old one-thread-per-group loop summed 240
  // shared-memory values back to back: a ~4 us serial tail on every CTA.
  const int cpg = C / groups;
  float* part = stats + (static_cast<size_t>(img) * kGroupNormMaxChunks + blockIdx.x) * groups * 2;
  {
    const int o = threadIdx.x >> 3, sub = threadIdx.x & 7;       // o = g * 2 + (0: sum, 1: sum of squares)
    float acc = 0.f;
    if (o < 2 * groups) {
      const int g = o >> 1, which = o & 1;
      const int n = ppp * cpg;
      for (int k = sub; k < n; k += 8) {
        const int l = k / cpg, c = g * cpg + k % cpg;
        acc += sm[(l * 2 + which) * C + c];
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (o < 2 * groups && sub == 0) part[o] = acc;
  }
  // the last CTA of this image to finish folds the partials (in chunk order) into mean / rstd
  __threadfence();
  __syncthreads();
  int* ticket = reinterpret_cast<int*>(gn_final(stats, n_img, groups) + static_cast<size_t>(n_img) * groups * 2);
  if (threadIdx.x == 0) s_last = atomicAdd(&ticket[img], 1) == static_cast<int>(gridDim.x) - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float inv_n = 1.0f / (static_cast<float>(hw) * cpg);
  float* fin = gn_final(stats, n_img, groups) + static_cast<size_t>(img) * groups * 2;
  const float* src = stats + static_cast<size_t>(img) * kGroupNormMaxChunks * groups * 2;
  {
    const int o = threadIdx.x >> 3, sub = threadIdx.x & 7;
    float acc = 0.f;
    if (o < 2 * groups)
      for (int ch = sub; ch < static_cast<int>(gridDim.x); ch += 8) acc += __ldcg(src + static_cast<size_t>(ch) * groups * 2 + o);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    const float other = __shfl_xor_sync(0xffffffffu, acc, 8);    // lanes of o ^ 1: (sum, sumsq) pair up
    if (o < 2 * groups && sub == 0 && (o & 1) == 0) {
      const float mean = acc * inv_n;
      const float var = fmaxf(other * inv_n - mean * mean, 0.f);
      fin[o] = mean;
      fin[o + 1] = rsqrtf(var + eps);
    }
  }
  if (threadIdx.x == 0) ticket[img] = 0;         // ready for the next launch / graph replay
}

// Statistics from the producing GEMMs' epilogue partials ([row block of 32][channel][sum, sumsq]) instead of
// a pass over the tensor.  grid (groups, n_img), 128 threads; fixed assignment + fixed reduction tree.
__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float* __restrict__ part0, int c0, const float* __restrict__ part1, int c1, int hw,
                   int groups, float eps, float* __restrict__ stats) {
  __shared__ float s_a[4], s_q[4];
  const int C = c0 + c1, cpg = C / groups;
  const int g = blockIdx.x, img = blockIdx.y, n_img = gridDim.y;
  const int nb = hw >> 5;                                  // 32-row blocks per image
  const int items = nb * cpg;
  auto load = [&](int k) {
    const int b = k / cpg, c = g * cpg + k % cpg;
    const size_t blk = static_cast<size_t>(img) * nb + b;
    return c < c0 ? __ldg(reinterpret_cast<const float2*>(part0 + (blk * c0 + c) * 2))
                  : __ldg(reinterpret_cast<const float2*>(part1 + (blk * c1 + (c - c0)) * 2));
  };
  // Eight loads in flight per thread (the kernel is pure latency: 1280 partials per block at 64x64 / C = 320);
  // fixed assignment and a fixed summation order keep the statistics bit-reproducible.
  float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int k = threadIdx.x;
  for (; k + 7 * 128 < items; k += 8 * 128) {
    float2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = load(k + j * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a8[j] += v[j].x; q8[j] += v[j].y; }
  }
  for (; k < items; k += 128) {
    const float2 v = load(k);
    a8[0] += v.x;
    q8[0] += v.y;
  }
  float a = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
  float q = ((q8[0] + q8[1]) + (q8[2] + q8[3])) + ((q8[4] + q8[5]) + (q8[6] + q8[7]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_q[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float ta = (s_a[0] + s_a[1]) + (s_a[2] + s_a[3]), tq = (s_q[0] + s_q[1]) + (s_q[2] + s_q[3]);
    const float inv_n = 1.0f / (static_cast<float>(hw) * cpg);
    const float mean = ta * inv_n;
    const float var = fmaxf(tq * inv_n - mean * mean, 0.f);
    float* fin = gn_final(stats, n_img, groups) + (static_cast<size_t>(img) * groups + g) * 2;
    fin[0] = mean;
    fin[1] = rsqrtf(var + eps);
  }
}

// grid (chunks, n_img): y = act((x - mean) * rstd * gamma + beta); scale / shift in registers.
__global__ void __launch_bounds__(kGnThreads, 2)
gn_apply_kernel(GnSrc s, int hw, int groups, float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, int silu, __nv_bfloat16* __restrict__ y) {
  const int C = s.c0 + s.c1;
  const int cpg = C / groups;
  const int vpp = C / 8;
  const int ppp = kGnThreads / vpp;
  const int img = blockIdx.y;
  const int v = threadIdx.x % vpp;
  const int pl = threadIdx.x / vpp;
  if (pl >= ppp) return;
  const float* fin = gn_final(stats, gridDim.y, groups) + static_cast<size_t>(img) * groups * 2;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    sc[j] = fin[g * 2 + 1] * __ldg(gamma + c);
    sh[j] = __ldg(beta + c) - fin[g * 2] * sc[j];
  }
  const int chunk = (hw + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * chunk;
  const int p_end = min(hw, p_begin + chunk);
  const size_t base = static_cast<size_t>(img) * hw;
  auto norm_store = [&](const uint4& u, size_t pix) {
    uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo = fmaf(bf16_lo(w[j]), sc[2 * j], sh[2 * j]);
      float hi = fmaf(bf16_hi(w[j]), sc[2 * j + 1], sh[2 * j + 1]);
      if (silu) {            // x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE MUFU op per element
        lo = silu_tanh(lo);  // (the exp + divide form needs two and made this kernel XU-bound: 48% XU busy)
        hi = silu_tanh(hi);
      }
      w[j] = pack_bf16(lo, hi);
    }
    *reinterpret_cast<uint4*>(y + pix * C + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  };
  for (int p = p_begin + pl; p < p_end; p += 8 * ppp) {
    uint4 u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      u[k] = p + k * ppp < p_end ? gn_load(s, base + p + k * ppp, v * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (p + k * ppp < p_end) norm_store(u[k], base + p + k * ppp);
  }
}

// ---- one-launch form for statistics that come from GEMM epilogue partials ------------------------------------
// grid (R, n_img) launched as thread-block clusters of R CTAs (one cluster per image, R <= 16): the finalize step
// of the two-kernel form (a 1024-CTA latency-bound launch plus a launch gap in front of every apply pass) becomes
// the prologue of the apply kernel.  CTA r of a cluster folds the row blocks [r nb / R, (r + 1) nb / R) of the
// partials into per-channel and then per-group sums in its own shared memory, the cluster barrier publishes them,
// every CTA adds the R group partials in rank order through distributed shared memory (fixed assignment, fixed
// order: bit-reproducible, no atomics, no global scratch) and then normalises its pixel chunk as gn_apply_kernel
// does.  The hardware co-schedules a cluster, so the barrier cannot deadlock whatever else is resident.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_dsmem(const float* local, uint32_t rank) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local)), r;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
  return v;
}

__global__ void __launch_bounds__(kGnThreads, 2)
gn_cluster_kernel(GnSrc s, const float* __restrict__ part0, const float* __restrict__ part1, int hw, int groups,
                  float eps, const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                  __nv_bfloat16* __restrict__ y) {
  extern __shared__ float sm_ch[];               // [2][C] per-channel (sum, sumsq) of this CTA's row blocks
  __shared__ float s_grp[64];                    // [groups][2] this CTA's group partials (read by the whole cluster)
  __shared__ float s_fin[64];                    // [groups][2] mean, rstd
  const int C = s.c0 + s.c1;
  const int cpg = C / groups;
  const int img = blockIdx.y;
  const int R = gridDim.x, rank = blockIdx.x;    // the cluster spans the x dimension
  const int nb = hw >> 5;                        // 32-row blocks per image
  const int b0 = static_cast<int>(static_cast<long long>(rank) * nb / R);
  const int b1 = static_cast<int>(static_cast<long long>(rank + 1) * nb / R);
  pdl_wait();                                    // SONIC_PDL: launched while the producing GEMM drains
  pdl_launch_dependents_small();
  for (int c = threadIdx.x; c < C; c += kGnThreads) {
    const float2* src = c < s.c0
        ? reinterpret_cast<const float2*>(part0) + static_cast<size_t>(img) * nb * s.c0 + c
        : reinterpret_cast<const float2*>(part1) + static_cast<size_t>(img) * nb * s.c1 + (c - s.c0);
    const int pitch = c < s.c0 ? s.c0 : s.c1;
    float a = 0.f, q = 0.f;
    int b = b0;
    for (; b + 8 <= b1; b += 8) {                // eight loads in flight; summed in block order
      float2 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + static_cast<size_t>(b + j) * pitch);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a += v[j].x; q += v[j].y; }
    }
    for (; b < b1; ++b) {
      const float2 v = __ldcg(src + static_cast<size_t>(b) * pitch);
      a += v.x;
      q += v.y;
    }
    sm_ch[c] = a;
    sm_ch[C + c] = q;
  }
  __syncthreads();
  {
    const int o = threadIdx.x >> 3, sub = threadIdx.x & 7;       // o = g * 2 + (0: sum, 1: sum of squares)
    float acc = 0.f;
    if (o < 2 * groups) {
      const float* src = sm_ch + (o & 1) * C + (o >> 1) * cpg;
      for (int k = sub; k < cpg; k += 8) acc += src[k];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (o < 2 * groups && sub == 0) s_grp[o] = acc;
  }
  cluster_arrive();
  // while the peers finish: the first batch of this thread's pixels is already on its way from HBM
  const int vpp = C / 8;
  const int ppp = kGnThreads / vpp;
  const int v = threadIdx.x % vpp;
  const int pl = threadIdx.x / vpp;
  const int chunk = (hw + R - 1) / R;
  const int p_begin = rank * chunk;
  const int p_end = min(hw, p_begin + chunk);
  const size_t base = static_cast<size_t>(img) * hw;
  uint4 u[8];
  if (pl < ppp) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      u[k] = p_begin + pl + k * ppp < p_end ? gn_load_cg(s, base + p_begin + pl + k * ppp, v * 8) : make_uint4(0, 0, 0, 0);
  }
  cluster_wait();
  if (threadIdx.x < 2 * groups) {
    float tot = 0.f;
    for (int r = 0; r < R; ++r) tot += ld_dsmem(&s_grp[threadIdx.x], static_cast<uint32_t>(r));
    s_fin[threadIdx.x] = tot;
  }
  cluster_arrive();                              // my remote reads are done (matched by the wait before exit)
  __syncthreads();
  float mean = 0.f, rstd = 0.f;
  if (threadIdx.x < groups) {
    const float inv_n = 1.0f / (static_cast<float>(hw) * cpg);
    mean = s_fin[2 * threadIdx.x] * inv_n;
    rstd = rsqrtf(fmaxf(s_fin[2 * threadIdx.x + 1] * inv_n - mean * mean, 0.f) + eps);
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    s_fin[2 * threadIdx.x] = mean;
    s_fin[2 * threadIdx.x + 1] = rstd;
  }
  __syncthreads();
  if (pl < ppp) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const int g = c / cpg;
      sc[j] = s_fin[g * 2 + 1] * __ldg(gamma + c);
      sh[j] = __ldg(beta + c) - s_fin[g * 2] * sc[j];
    }
    auto norm_store = [&](const uint4& u, size_t pix) {
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo = fmaf(bf16_lo(w[j]), sc[2 * j], sh[2 * j]);
        float hi = fmaf(bf16_hi(w[j]), sc[2 * j + 1], sh[2 * j + 1]);
        if (silu) {
          lo = silu_tanh(lo);
          hi = silu_tanh(hi);
        }
        w[j] = pack_bf16(lo, hi);
      }
      *reinterpret_cast<uint4*>(y + pix * C + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    };
    for (int p = p_begin + pl; p < p_end; p += 8 * ppp) {
      if (p != p_begin + pl) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          u[k] = p + k * ppp < p_end ? gn_load_cg(s, base + p + k * ppp, v * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (p + k * ppp < p_end) norm_store(u[k], base + p + k * ppp);
    }
  }
  cluster_wait();                                // no CTA leaves while a peer may still read its s_grp
}

// One warp per kRows rows (all loads of the warp's rows are issued before the first reduction, so enough
// bytes are in flight per SM to cover HBM latency); rows live in registers between the two passes.
template <int kVecPerLane, int kRows>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int rows, int C, float eps,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y) {
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRows;
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  const int nvec = C / 8;
  uint4 u[kRows][kVecPerLane];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(min(row0 + r, rows - 1)) * C);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int v = lane + i * 32;
      u[r][i] = v < nvec ? __ldg(xr + v) : make_uint4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    if (row0 + r >= rows) break;
    float f[kVecPerLane][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const uint32_t w[4] = {u[r][i].x, u[r][i].y, u[r][i].z, u[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[i][2 * j] = bf16_lo(w[j]);
        f[i][2 * j + 1] = bf16_hi(w[j]);
        sum += f[i][2 * j] + f[i][2 * j + 1];       // padding vectors are zero
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
    uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row0 + r) * C);
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
}

// LayerNorm for C = 40 * kLanes (320 / 640 / 1280: every LayerNorm of the SD UNet): kLanes lanes share a row,
// five 16-byte vectors each, so all 32 lanes carry data (the one-row-per-warp kernel idles 24 lanes on its
// second vector at C=320) and a row reduction is log2(kLanes) shuffles.  Each warp walks kGroups row groups
// with the NEXT group's loads in flight; gamma / beta live in registers for the whole walk.
template <int kLanes>
__global__ void __launch_bounds__(128, 3)
layernorm40_kernel(const __nv_bfloat16* __restrict__ x, int rows, float eps, const float* __restrict__ gamma,
                   const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int groups_per_warp) {
  constexpr int kRowsPerGroup = 32 / kLanes;
  constexpr int C = 40 * kLanes;
  const int lane = threadIdx.x & 31;
  const int sub = lane / kLanes;                 // row inside the group
  const int l = lane % kLanes;                   // position inside the row
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int group0 = warp_global * groups_per_warp;
  float g[5][8], b[5][8];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int v = l + i * kLanes;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
    g[i][0] = g0.x; g[i][1] = g0.y; g[i][2] = g0.z; g[i][3] = g0.w; g[i][4] = g1.x; g[i][5] = g1.y; g[i][6] = g1.z; g[i][7] = g1.w;
    b[i][0] = b0.x; b[i][1] = b0.y; b[i][2] = b0.z; b[i][3] = b0.w; b[i][4] = b1.x; b[i][5] = b1.y; b[i][6] = b1.z; b[i][7] = b1.w;
  }
  auto load = [&](int grp, uint4 (&u)[5]) {
    const int row = min((group0 + grp) * kRowsPerGroup + sub, rows - 1);
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C);
#pragma unroll
    for (int i = 0; i < 5; ++i) u[i] = __ldg(xr + l + i * kLanes);
  };
  uint4 cur[5], nxt[5];
  load(0, cur);
  for (int grp = 0; grp < groups_per_warp; ++grp) {
    const int row = (group0 + grp) * kRowsPerGroup + sub;
    if ((group0 + grp) * kRowsPerGroup >= rows) break;
    if (grp + 1 < groups_per_warp) load(grp + 1, nxt);
    float f[5][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[i][2 * j] = bf16_lo(w[j]);
        f[i][2 * j + 1] = bf16_hi(w[j]);
        sum += f[i][2 * j] + f[i][2 * j + 1];
      }
    }
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / C) + eps);
    if (row < rows) {
      uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * C);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = pack_bf16((f[i][2 * j] - mean) * rstd * g[i][2 * j] + b[i][2 * j],
                           (f[i][2 * j + 1] - mean) * rstd * g[i][2 * j + 1] + b[i][2 * j + 1]);
        yr[l + i * kLanes] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) cur[i] = nxt[i];
  }
}

// Folded LayerNorm, between producer and consumer GEMM (gemm.cuh): per row, fold the producer's (sum, sumsq)
// partials in a fixed order and emit (i) rstd, the consumer's per-row epilogue scale, and (ii) the 16 bytes of the
// side tensor row that carry -mean and std = 1 / rstd through the tensor core as bf16 hi / lo pairs:
//   side[row][0..7] = (-mu_hi, -mu_hi, -mu_lo, -mu_lo, std_hi, std_hi, std_lo, std_lo)   (columns 8..63 stay zero)
// against B columns (s_hi, s_lo, s_hi, s_lo, b_hi, b_lo, b_hi, b_lo): the four products of each hi / lo pair.
__global__ void __launch_bounds__(256)
ln_side_kernel(const float2* __restrict__ part, int parts, int M, float inv_k, float eps,
               __nv_bfloat16* __restrict__ side, float* __restrict__ rstd) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();                                    // SONIC_PDL: before the first global access and before any exit
  pdl_launch_dependents_small();
  if (row >= M) return;
  const float2* st = part + static_cast<size_t>(row) * parts;
  float sa = 0.f, sq = 0.f;
  for (int i = 0; i < parts; ++i) {
    const float2 v = __ldcg(st + i);             // written by the GEMM this kernel may overlap (SONIC_PDL): through L2
    sa += v.x;
    sq += v.y;
  }
  const float mean = sa * inv_k;
  const float var = fmaxf(sq * inv_k - mean * mean, 0.f);
  const float r = rsqrtf(var + eps);
  const float sd = (var + eps) * r;
  const float m = -mean;
  const __nv_bfloat16 m_hi = __float2bfloat16_rn(m), d_hi = __float2bfloat16_rn(sd);
  const __nv_bfloat16 m_lo = __float2bfloat16_rn(m - __bfloat162float(m_hi));
  const __nv_bfloat16 d_lo = __float2bfloat16_rn(sd - __bfloat162float(d_hi));
  auto pair = [](__nv_bfloat16 a, __nv_bfloat16 b) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
  };
  *reinterpret_cast<uint4*>(side + static_cast<size_t>(row) * 64) =
      make_uint4(pair(m_hi, m_hi), pair(m_lo, m_lo), pair(d_hi, d_hi), pair(d_lo, d_lo));
  rstd[row] = r;
}

}  // namespace

int ln_side_launch(const float* partials, int parts, int M, int K, float eps, void* side, float* rstd,
                   cudaStream_t stream) {
  SONIC_REQUIRE(partials && side && rstd && parts > 0 && M > 0 && K > 0, "ln_side: bad argument");
  if (pdl_enabled())
    SONIC_CUDA(launch_kernel_ex(ln_side_kernel, dim3((M + 255) / 256), dim3(256), 0, stream, 1,
                                reinterpret_cast<const float2*>(partials), parts, M, 1.0f / static_cast<float>(K), eps,
                                static_cast<__nv_bfloat16*>(side), rstd));
  else
    ln_side_kernel<<<(M + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float2*>(partials), parts, M,
                                                       1.0f / static_cast<float>(K), eps,
                                                       static_cast<__nv_bfloat16*>(side), rstd);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

// CTAs per cluster of the one-launch form for this operator, or 0 if it runs as finalize (or statistics) + apply.
static int gn_cluster_size(const GroupNormOp& op) {
  if (!op.part0) return 0;
  const int C = op.c0 + (op.x1 ? op.c1 : 0);
  const int ppp = kGnThreads / (C / 8);
  static const int fused_mode = [] { const char* e = getenv("SONIC_GN_CLUSTER"); return e ? atoi(e) : 1; }();
  const int R = std::min(std::min(8, (2 * 148) / std::max(1, op.n_img)), std::max(1, op.hw / (4 * ppp)));
  const bool small = static_cast<size_t>(op.n_img) * op.hw * C * 2 < (8u << 20);      // latency-bound either way
  return (fused_mode && op.groups <= 32 && (R * op.n_img >= 200 || small)) ? R : 0;
}

int groupnorm_launch_count(const GroupNormOp& op) { return gn_cluster_size(op) > 0 ? 1 : 2; }

int groupnorm_launch(const GroupNormOp& op, cudaStream_t stream) {
  const int C = op.c0 + (op.x1 ? op.c1 : 0);
  SONIC_REQUIRE(op.x0 && op.y && op.stats && op.gamma && op.beta, "groupnorm: null operand");
  SONIC_REQUIRE(C % op.groups == 0 && op.c0 % 8 == 0 && C % 8 == 0, "groupnorm: C=%d groups=%d unsupported",
                C, op.groups);
  SONIC_REQUIRE(C / 8 <= kGnThreads, "groupnorm: C=%d too wide (max %d)", C, 8 * kGnThreads);
  GnSrc s{static_cast<const __nv_bfloat16*>(op.x0), static_cast<const __nv_bfloat16*>(op.x1), op.c0,
          op.x1 ? op.c1 : 0, op.ld0 ? op.ld0 : op.c0, op.ld1 ? op.ld1 : op.c1};
  const int ppp = kGnThreads / (C / 8);
  // >= 4 pixels per pixel-lane per CTA; ONE wave of 3 resident CTAs per SM: fat CTAs amortise the
  // per-CTA reduction / ticket tail (37 thin chunks per image ran the statistics pass at 27% of HBM peak)
  static const int slots = [] { const char* e = getenv("SONIC_GN_SLOTS"); return e ? atoi(e) : 3 * 148; }();
  int chunks = std::max(1, std::min(op.hw / (4 * ppp), slots / std::max(1, op.n_img)));
  chunks = std::min(chunks, kGroupNormMaxChunks);
  dim3 grid(chunks, op.n_img);
  const size_t smem = static_cast<size_t>(ppp) * 2 * C * sizeof(float);
  SONIC_REQUIRE(smem <= 48 * 1024, "groupnorm: C=%d needs too much shared memory", C);
  if (op.part0) {
    SONIC_REQUIRE(op.hw % 32 == 0, "groupnorm: fused statistics need hw %% 32 == 0 (got %d)", op.hw);
    SONIC_REQUIRE(s.ld0 == s.c0 && (!op.x1 || s.ld1 == s.c1), "groupnorm: fused statistics need dense rows");
    // One launch: a cluster of R <= 8 CTAs per image (gn_cluster_kernel).  R: one wave of two CTAs per SM and at least
    // four pixels per pixel lane and CTA.  Measured at UNet batch 32: 2.05 -> 1.73 ms per step over the 61 GroupNorms
    // (64x64x320 + SiLU: 48.7 -> 40.1 us); clusters of 9 CTAs place badly (54.8 us), and 16 x 16 CTAs on the VAE
    // decoder's 1 GB tensors lose to the 432 CTAs of the two-kernel form (42.2 vs 40.7 ms per decode), so large
    // tensors of small batches keep the finalize + apply pair.  SONIC_GN_CLUSTER=0 disables this path (A/B).
    const int R = gn_cluster_size(op);
    if (R > 0) {
      // R == 1 still goes through the cluster attribute (a cluster of one), as before
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(R, op.n_img);
      cfg.blockDim = dim3(kGnThreads);
      cfg.dynamicSmemBytes = static_cast<size_t>(2 * C) * sizeof(float);
      cfg.stream = stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = R;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.numAttrs = 1;
      if (pdl_enabled()) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
      }
      cfg.attrs = attr;
      SONIC_CUDA(cudaLaunchKernelEx(&cfg, gn_cluster_kernel, s, op.part0, op.part1, op.hw, op.groups, op.eps, op.gamma,
                                    op.beta, op.silu, static_cast<__nv_bfloat16*>(op.y)));
      return 0;
    }
    gn_finalize_kernel<<<dim3(op.groups, op.n_img), 128, 0, stream>>>(op.part0, s.c0, op.part1, s.c1, op.hw, op.groups,
                                                                      op.eps, op.stats);
  } else {
    gn_stats_kernel<<<grid, kGnThreads, smem, stream>>>(s, op.hw, op.groups, op.eps, op.stats);
  }
  gn_apply_kernel<<<grid, kGnThreads, 0, stream>>>(s, op.hw, op.groups, op.stats, op.gamma, op.beta, op.silu,
                                                   static_cast<__nv_bfloat16*>(op.y));
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_launch(const void* x, void* y, int rows, int C, float eps, const float* gamma,
                     const float* beta, cudaStream_t stream) {
  SONIC_REQUIRE(C % 8 == 0 && C <= 2048, "layernorm: C=%d unsupported", C);
  const int nvec = C / 8;
  const int vpl = (nvec + 31) / 32;
  const int warps = 8;
  auto xb = static_cast<const __nv_bfloat16*>(x);
  auto yb = static_cast<__nv_bfloat16*>(y);
  if (C == 320 || C == 640 || C == 1280) {
    const int lanes = C / 40, rows_per_group = 32 / lanes;
    const int n_groups = (rows + rows_per_group - 1) / rows_per_group;
    // ~4 resident waves of warps (3 CTAs x 4 warps per SM), at least 2 groups per warp
    const int w40 = 4;
    const int gpw = std::max(2, (n_groups + 4 * 148 * 3 * w40 - 1) / (4 * 148 * 3 * w40));
    const int n_warps = (n_groups + gpw - 1) / gpw;
    dim3 grid((n_warps + w40 - 1) / w40);
    if (lanes == 8) layernorm40_kernel<8><<<grid, w40 * 32, 0, stream>>>(xb, rows, eps, gamma, beta, yb, gpw);
    else if (lanes == 16) layernorm40_kernel<16><<<grid, w40 * 32, 0, stream>>>(xb, rows, eps, gamma, beta, yb, gpw);
    else layernorm40_kernel<32><<<grid, w40 * 32, 0, stream>>>(xb, rows, eps, gamma, beta, yb, gpw);
    SONIC_CUDA(cudaGetLastError());
    return 0;
  }
#define SONIC_LN(V, R)                                                                                     \
  layernorm_kernel<V, R><<<dim3((rows + warps * R - 1) / (warps * R)), warps * 32, 0, stream>>>(xb, rows, C, eps, \
                                                                                                gamma, beta, yb)
  switch (vpl) {
    case 1: SONIC_LN(1, 4); break;
    case 2: SONIC_LN(2, 4); break;
    case 3: SONIC_LN(3, 2); break;
    case 4: SONIC_LN(4, 2); break;
    case 5: SONIC_LN(5, 2); break;
    default: SONIC_LN(8, 1); break;
  }
#undef SONIC_LN
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sonic
