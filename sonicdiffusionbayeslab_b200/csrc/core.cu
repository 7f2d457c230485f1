// Error state and TMA descriptor encoding shared by every operator in libsonic.
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace sonic {

namespace {
thread_local char g_err[1024] = "";
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SONIC_PDL"); return e && atoi(e) != 0; }();
  return on;
}

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%d) at %s:%d: %s", cudaGetErrorString(e), static_cast<int>(e), file,
            line, what);
  return -1;
}

int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    SONIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    SONIC_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess,
                  "cuTensorMapEncodeTiled is not available from this driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  SONIC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16B aligned",
                base);
  cuuint64_t d[5];
  cuuint64_t s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    SONIC_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    s[i] = strides_bytes[i];
    SONIC_REQUIRE(s[i] % 16 == 0, "tensor map stride[%d]=%llu not a multiple of 16", i,
                  static_cast<unsigned long long>(s[i]));
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SONIC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)",
                static_cast<int>(r), rank);
  return 0;
}

}  // namespace sonic
