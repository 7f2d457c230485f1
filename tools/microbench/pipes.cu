// Micro-benchmarks that decide kernel design choices on B200 (run under gpurun):
//   1. MUFU ex2 throughput: f32 vs packed f16x2 vs packed bf16x2   (softmax of the attention kernel)
//   2. FFMA vs packed FFMA2 (fma.rn.f32x2) throughput               (softmax scale/sub, GELU polynomial)
//   3. L2 -> SM fill bandwidth with cp.async.bulk, all SMs pulling  (operand feed of the implicit GEMM)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <int MODE>
__global__ void __launch_bounds__(256) mufu_kernel(float* out, long long* cyc) {
  uint32_t r[kChains];
  float f[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) { r[i] = 0x34003400u + threadIdx.x + i; f[i] = -0.001f * (threadIdx.x + i); }
  long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
      if (MODE == 4) {
        uint64_t v = (static_cast<uint64_t>(__float_as_uint(f[i])) << 32) | r[i];
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        f[i] = __uint_as_float(static_cast<uint32_t>(v >> 32)); r[i] = static_cast<uint32_t>(v);
      }
      if (MODE == 5) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += f[i] + __uint_as_float(r[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run_mufu(const char* name, int elems_per_op, float* out, long long* cyc, int ctas_per_sm) {
  int n_cta = 148 * ctas_per_sm;
  mufu_kernel<MODE><<<n_cta, 256>>>(out, cyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  mufu_kernel<MODE><<<n_cta, 256>>>(out, cyc);
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[148 * 8]; CK(cudaMemcpy(h, cyc, sizeof(long long) * n_cta, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < n_cta; ++i) avg += h[i]; avg /= n_cta;
  double ops = static_cast<double>(kIters) * kChains * 256 * ctas_per_sm;   // thread-ops per SM
  printf("%-28s ctas/SM=%d  %.1f thread-ops/clk/SM  = %.1f elems/clk/SM   (%.3f ms, %.0f cyc)\n", name, ctas_per_sm,
         ops / avg, ops * elems_per_op / avg, ms, avg);
  return 0;
}

// ---- L2 -> SM bulk-copy bandwidth
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void __launch_bounds__(128) l2_fill_kernel(const uint8_t* src, size_t region, int chunk, int n_chunks, int rounds,
                                                      long long* cyc, int same) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* base = src + (same ? 0 : static_cast<size_t>(blockIdx.x) * region);
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int issued = 0, done = 0;
    const int total = n_chunks * rounds;
    while (done < total) {
      while (issued < total && issued - done < 4) {
        const int st = issued & 3;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[st])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(sm + st * chunk)), "l"(base + static_cast<size_t>(issued % n_chunks) * chunk), "r"(chunk),
                     "r"(s32(&bar[st])) : "memory");
        ++issued;
      }
      const int st = done & 3;
      const uint32_t parity = (done >> 2) & 1;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(s32(&bar[st])), "r"(parity) : "memory");
      ++done;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int run_l2(const uint8_t* buf, long long* cyc, int chunk, size_t region, int same, int grid) {
  const int n_chunks = static_cast<int>(region / chunk), rounds = 64;
  cudaFuncSetAttribute(l2_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * chunk);
  for (int rep = 0; rep < 2; ++rep) {
    l2_fill_kernel<<<grid, 128, 4 * chunk>>>(buf, region, chunk, n_chunks, rounds, cyc, same);
    CK(cudaDeviceSynchronize());
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  l2_fill_kernel<<<grid, 128, 4 * chunk>>>(buf, region, chunk, n_chunks, rounds, cyc, same);
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double bytes = static_cast<double>(chunk) * n_chunks * rounds;
  printf("L2 fill: grid=%3d chunk=%5d B region/CTA=%6zu KB same=%d : %.1f B/clk/SM, chip %.2f TB/s (%.3f ms)\n", grid, chunk,
         region >> 10, same, bytes / avg, bytes * grid / (ms * 1e-3) / 1e12, ms);
  return 0;
}

int main() {
  float* out; long long* cyc; uint8_t* buf;
  CK(cudaMalloc(&out, sizeof(float) * 148 * 8 * 256));
  CK(cudaMalloc(&cyc, sizeof(long long) * 148 * 8));
  const size_t buf_bytes = 148ull * 512 * 1024;
  CK(cudaMalloc(&buf, buf_bytes));
  CK(cudaMemset(buf, 1, buf_bytes));
  for (int c : {1, 2, 4}) {
    if (run_mufu<0>("ex2.approx.ftz.f32", 1, out, cyc, c)) return 1;
    if (run_mufu<1>("ex2.approx.f16x2", 2, out, cyc, c)) return 1;
    if (run_mufu<2>("ex2.approx.ftz.bf16x2", 2, out, cyc, c)) return 1;
    if (run_mufu<3>("fma.rn.f32", 1, out, cyc, c)) return 1;
    if (run_mufu<4>("fma.rn.f32x2", 2, out, cyc, c)) return 1;
    if (run_mufu<5>("tanh.approx.f32", 1, out, cyc, c)) return 1;
  }
  for (int grid : {1, 32, 148}) {
    run_l2(buf, cyc, 16384, 256 * 1024, 0, grid);
    run_l2(buf, cyc, 32768, 256 * 1024, 0, grid);
  }
  run_l2(buf, cyc, 32768, 256 * 1024, 1, 148);
  run_l2(buf, cyc, 32768, 512 * 1024, 0, 148);
  return 0;
}
