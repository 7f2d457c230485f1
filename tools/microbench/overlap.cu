// Do the MUFU-bound softmax loop (tcgen05.ld -> ex2 -> tcgen05.st) and small tcgen05.mma chains overlap on one SM?
// Each CTA: warps 0-3 run the softmax loop on their TMEM lane quarter, warp 4 issues MMAs (SS N=64 or TS N=48).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o overlap overlap.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// mma_mode: 0 none, 1 SS (A,B smem) N=64, 2 TS (A tmem) N=48
__global__ void __launch_bounds__(160) k(int sm_iters, int n_mma, int mma_mode, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 160) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  long long t0 = clock64(), t1 = t0;
  if (warp < 4) {
    const uint32_t base = tm + (static_cast<uint32_t>(warp * 32) << 16);       // columns [0,32) S, P over [0,16)
    float psm[4] = {0, 0, 0, 0};
    const float scale = 1e-30f, msc = 1e-3f * threadIdx.x;
    uint32_t v[32];
    for (int it = 0; it < sm_iters; ++it) {
      ld32(base, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v[i]), scale, -msc)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v[i + 1]), scale, -msc)));
        psm[(i >> 1) & 3] += e0 + e1;
        __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
      st16(base + 32, pk);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    t1 = clock64();
    sink[blockIdx.x * 160 + threadIdx.x] = psm[0] + psm[1] + psm[2] + psm[3];
  } else if (mma_mode != 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
    const int N = mma_mode == 1 ? 64 : 48;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
    const uint64_t da = desc128(s32(sm), 16, 1024), db = desc128(s32(sm) + 16384, 16, 1024);
    for (int i = 0; i < n_mma; i += 4) {
      if (leader) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (mma_mode == 1)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm + 64),
                         "l"(da + u * 2), "l"(db + u * 2), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm + 64),
                         "r"(tm + 48 + (u & 1) * 8), "l"(db + u * 2), "r"(idesc), "r"(1u) : "memory");
        }
      }
      __syncwarp();
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp <= 4) out[blockIdx.x * 5 + warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(slot) : "memory");
}
int main() {
  long long* out; float* sink;
  const int cps = 4, grid = 148 * cps;
  CK(cudaMalloc(&out, grid * 5 * 8)); CK(cudaMalloc(&sink, grid * 160 * 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
  const int sm_iters = 512;
  printf("4 CTAs/SM; softmax loop = %d x (ld32, 32 ex2, st16) per warp; MUFU floor per SM = %d cycles\n", sm_iters,
         sm_iters * 32 * 8 * cps);
  for (int mode : {0, 1, 2}) {
    for (int n_mma : {0, 512, 1024, 2048}) {
      if ((mode == 0) != (n_mma == 0)) continue;
      for (int si : {0, sm_iters}) {
        if (si == 0 && n_mma == 0) continue;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<<<grid, 160, 40 * 1024>>>(si, n_mma, mode, out, sink);
        cudaEventRecord(a);
        k<<<grid, 160, 40 * 1024>>>(si, n_mma, mode, out, sink);
        cudaEventRecord(b);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b);
        static long long h[148 * 4 * 5];
        CK(cudaMemcpy(h, out, sizeof(long long) * grid * 5, cudaMemcpyDeviceToHost));
        double tsm = 0, tmma = 0;
        for (int c = 0; c < grid; ++c) { tsm += h[c * 5]; tmma += h[c * 5 + 4]; }
        printf("mma_mode %d n_mma/CTA %4d softmax_iters %3d : kernel %.1f us | softmax warp %.0f cyc | mma warp %.0f cyc\n", mode,
               n_mma, si, ms * 1e3, tsm / grid, tmma / grid);
      }
    }
  }
  return 0;
}
