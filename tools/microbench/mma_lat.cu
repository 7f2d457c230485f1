// tcgen05.mma issue / completion latency for chains of small MMAs (one CTA per SM, one issuing lane).
//   dependent chain: n MMAs accumulate into the same TMEM tile; independent: alternate two accumulators.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_lat mma_lat.cu
#include <cuda_runtime.h>
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
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) k(int n, int N, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
  if (warp == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
    const uint64_t da = desc128(s32(sm), 16, 1024), db = desc128(s32(sm) + 16384, 16, 1024);
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < n; i += 4) {
        if (leader) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (i + u < n) {
              const uint32_t d = tm + 64;
              const uint32_t acc = (i + u) > 1 || (!(mode & 1) && (i + u) > 0);
              if (mode & 2) mma_ts(d, tm + u * 8, db + u * 2, idesc, acc);
              else mma_ss(d, da + u * 2, db + u * 2, idesc, acc);
            }
          }
        }
        __syncwarp();
      }
      long long t1 = clock64();
      if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      __syncwarp();
      long long t2 = clock64();
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(s32(&bar)), "r"(rep & 1) : "memory");
      long long t3 = clock64();
      if (blockIdx.x == 0 && threadIdx.x == 0 && rep < 3) { out[rep * 3 + 0] = t1 - t0; out[rep * 3 + 1] = t2 - t0; out[rep * 3 + 2] = t3 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(slot) : "memory");
}
int main() {
  long long* out; CK(cudaMalloc(&out, 9 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  printf("mode: 0 SS dependent, 1 SS two accumulators, 2 TS dependent; cycles (3rd repetition): issue-done / after-commit / complete\n");
  for (int cps : {1, 2, 4})
  for (int mode : {0, 2})
    for (int N : {48, 64})
      for (int n : {32}) {
        printf("CTAs/SM=%d ", cps);
        k<<<148 * cps, 128, 48 * 1024>>>(n, N, mode, out);
        CK(cudaDeviceSynchronize());
        long long h[9]; CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        printf("mode %d N=%3d n=%2d : %5lld / %5lld / %5lld   (%.1f cycles per MMA, floor %d)\n", mode, N, n, h[6], h[7], h[8],
               (double)h[8] / n, N / 2);
      }
  return 0;
}
