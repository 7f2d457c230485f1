// CLIP image preprocessing on the GPU, arithmetic-identical to the PIL path the reference's metric runs on the
// host (/root/reference/src/metrics/metrics.py:25-41 -> torchmetrics CLIPScore -> HF CLIPImageProcessor):
//   [optional quantise (x * 255 -> uint8, base_experiment.py:198-199)] -> bicubic resize of the shortest edge to 224
//   with PIL's antialiasing two-pass resampler (Pillow Resample.c: horizontal pass, uint8 intermediate, vertical
//   pass, 22-bit fixed-point coefficients, round-half-up, clip to [0, 255]) -> centre crop -> * 1/255 -> (x - mean)
//   / std.
// ONE kernel: a CTA owns a tile of output rows of one channel of one image, runs the horizontal pass for the input
// rows that tile needs into shared memory (uint8, exactly PIL's intermediate image), then the vertical pass, and
// writes either NCHW fp32 or -- fused with the ViT patch cut -- bf16 rows of the patch-embedding GEMM's A operand.
// Coefficient tables come from the host (kernels.py builds them with the same double arithmetic as Pillow's
// precompute_coeffs / normalize_coeffs_8bpc), so the result is bit-identical to PIL's uint8 image.
#include "ops.cuh"

namespace sonic {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;        // Pillow: PRECISION_BITS
constexpr int kTileRows = 16;

__device__ __forceinline__ int clip8(int v) {     // Pillow: clip8_lookups[v >> PRECISION_BITS]
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

template <typename T> struct Px;
template <> struct Px<uint8_t> {
  static __device__ __forceinline__ int get(const uint8_t* p) { return *p; }
};
template <> struct Px<float> {                     // (x * 255).to(uint8): truncation toward zero, saturating
  static __device__ __forceinline__ int get(const float* p) {
    const float v = *p * 255.0f;
    return v <= 0.f ? 0 : (v >= 255.f ? 255 : static_cast<int>(v));
  }
};
template <> struct Px<__nv_bfloat16> {
  static __device__ __forceinline__ int get(const __nv_bfloat16* p) {
    const float v = round_bf16(__bfloat162float(*p) * 255.0f);   // a bf16 tensor times 255 is rounded to bf16 first
    return v <= 0.f ? 0 : (v >= 255.f ? 255 : static_cast<int>(v));
  }
};

struct PreParams {
  int n_img, H, W;              // input images, NCHW, 3 channels
  const int* hb; const int* hk; int hks;     // horizontal bounds [nw][2], coefficients [nw][hks]
  const int* vb; const int* vk; int vks;     // vertical   bounds [nh][2], coefficients [nh][vks]
  int top, left, S;             // centre crop offset inside the resized image, output size
  float mean[3], stdv[3];
  int out_mode;                 // 0: fp32 NCHW [n][3][S][S]   1: bf16 patch rows [n*G*G][3*P*P]
  int patch;
  int max_rows;                 // input rows one tile may need (shared-memory pitch)
};

template <typename T>
__global__ void __launch_bounds__(256)
clip_preprocess_kernel(const T* __restrict__ src, void* __restrict__ out, const PreParams p) {
  extern __shared__ uint8_t tile[];               // [max_rows][S] horizontally resampled rows (PIL's intermediate)
  const int oy0 = blockIdx.x * kTileRows;
  const int rows_out = min(kTileRows, p.S - oy0);
  const int c = blockIdx.y, img = blockIdx.z;
  const int r0 = p.vb[2 * (oy0 + p.top)];
  const int last = oy0 + rows_out - 1 + p.top;
  const int r1 = p.vb[2 * last] + p.vb[2 * last + 1];
  const int nrows = r1 - r0;
  const T* plane = src + (static_cast<size_t>(img) * 3 + c) * p.H * p.W;
  // ---- horizontal pass over the input rows this tile needs
  for (int i = threadIdx.x; i < nrows * p.S; i += blockDim.x) {
    const int r = i / p.S, ox = i % p.S;
    const int xr = ox + p.left;
    const int xmin = p.hb[2 * xr], cnt = p.hb[2 * xr + 1];
    const int* k = p.hk + static_cast<size_t>(xr) * p.hks;
    const T* row = plane + static_cast<size_t>(r0 + r) * p.W + xmin;
    int ss = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < cnt; ++x) ss += Px<T>::get(row + x) * __ldg(k + x);
    tile[r * p.S + ox] = static_cast<uint8_t>(clip8(ss));
  }
  __syncthreads();
  // ---- vertical pass + rescale + normalise
  const float mean = p.mean[c], stdv = p.stdv[c];
  for (int i = threadIdx.x; i < rows_out * p.S; i += blockDim.x) {
    const int oy = oy0 + i / p.S, ox = i % p.S;
    const int yr = oy + p.top;
    const int ymin = p.vb[2 * yr], cnt = p.vb[2 * yr + 1];
    const int* k = p.vk + static_cast<size_t>(yr) * p.vks;
    int ss = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < cnt; ++y) ss += tile[(ymin - r0 + y) * p.S + ox] * __ldg(k + y);
    const int u = clip8(ss);
    const float x = static_cast<float>(static_cast<double>(u) * (1.0 / 255.0));     // HF rescale: float64 product -> float32
    const float v = __fdiv_rn(__fsub_rn(x, mean), stdv);
    if (p.out_mode == 0) {
      static_cast<float*>(out)[((static_cast<size_t>(img) * 3 + c) * p.S + oy) * p.S + ox] = v;
    } else {
      const int G = p.S / p.patch, P = p.patch;
      const size_t row = (static_cast<size_t>(img) * G + oy / P) * G + ox / P;
      static_cast<__nv_bfloat16*>(out)[row * (3 * P * P) + (c * P + oy % P) * P + ox % P] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace

int clip_preprocess_launch(const void* images, int dtype, int n_img, int H, int W, const int* hb, const int* hk, int hks,
                           const int* vb, const int* vk, int vks, int max_rows, int top, int left, int S,
                           const float* mean3, const float* std3, void* out, int out_mode, int patch,
                           cudaStream_t stream) {
  SONIC_REQUIRE(images && out && hb && hk && vb && vk && mean3 && std3, "clip_preprocess: null argument");
  SONIC_REQUIRE(n_img > 0 && S > 0 && max_rows > 0, "clip_preprocess: bad shape");
  SONIC_REQUIRE(out_mode == 0 || (out_mode == 1 && patch > 0 && S % patch == 0), "clip_preprocess: bad output mode");
  PreParams p;
  p.n_img = n_img; p.H = H; p.W = W;
  p.hb = hb; p.hk = hk; p.hks = hks; p.vb = vb; p.vk = vk; p.vks = vks;
  p.top = top; p.left = left; p.S = S;
  for (int i = 0; i < 3; ++i) { p.mean[i] = mean3[i]; p.stdv[i] = std3[i]; }
  p.out_mode = out_mode; p.patch = patch; p.max_rows = max_rows;
  const size_t smem = static_cast<size_t>(max_rows) * S;
  SONIC_REQUIRE(smem <= 48 * 1024, "clip_preprocess: tile needs %zu bytes of shared memory", smem);
  dim3 grid((S + kTileRows - 1) / kTileRows, 3, n_img);
  if (dtype == 2)
    clip_preprocess_kernel<uint8_t><<<grid, 256, smem, stream>>>(static_cast<const uint8_t*>(images), out, p);
  else if (dtype == kF32)
    clip_preprocess_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(images), out, p);
  else if (dtype == kBF16)
    clip_preprocess_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(images), out, p);
  else
    SONIC_REQUIRE(false, "clip_preprocess: unknown dtype %d", dtype);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sonic
