// Implicit-GEMM operator (3x3 / 1x1 convolution and Linear) on tcgen05 tensor cores.
// One persistent, warp-specialised kernel serves every GEMM-shaped layer of the UNet:
//   out[M, N] = epilogue( sum_{tap, k} A[pixel(m) + offset(tap), k] * Wt[tap][n][k] )
// A is an NHWC bf16 activation (optionally the channel-concat of two tensors), fetched by
// TMA as shifted (tile_n x tile_h x tile_w) x 64-channel boxes -- out-of-image taps are
// zero-filled by the TMA unit, so no im2col buffer ever exists.
#pragma once
#include "common.cuh"

namespace sonic {

enum GemmEpilogue { kEpiNone = 0, kEpiGeglu = 1, kEpiQuickGelu = 2 };

enum GemmMode {
  kModeNormal = 0,
  kModeStride2 = 1,              // 3x3, stride 2, pad 1: the nine taps read four PARITY views of the input
  kModeUpsample = 2,             // 3x3 over a nearest-2x upsampled input, as four 2x2 convolutions (one per phase)
};

struct GemmParams {
  CUtensorMap tm_a0, tm_a1, tm_b;
  CUtensorMap tm_out, tm_res;    // [M][n_out] output / residual, 32x32 boxes (staged epilogue)
  CUtensorMap tm_x[3];           // mode 1: parity views (0,1) (1,0) (1,1) of A (tm_a0 is (0,0));
                                 // mode 2: output views of phases (0,1) (1,0) (1,1) (tm_out is phase (0,0))
  int mode;
  int M, N;                      // GEMM rows (= n_img*H*W) and B rows (= out channels before GEGLU)
  int k_chunks0, k_chunks1;      // 64-wide K chunks taken from source 0 / source 1
  int taps;                      // 1 or 9
  int H, W, n_img;
  int tile_w, tile_h, tile_n;    // tile_w*tile_h*tile_n == 128 rows of one M tile
  int tiles_w, tiles_h, tiles_img;
  int m_tiles, n_tiles, block_n, stages;
  int n_acc, acc_stride;         // TMEM accumulator ring: buffers and columns per buffer
  int m_tiles_src;               // mode 2: M tiles of ONE phase (m_tiles = 4 * m_tiles_src)
  uint32_t idesc;
  const float* bias;             // [N] or null
  const float* row_bias;         // [n_img][N] or null
  const __nv_bfloat16* residual; // [M][ld_res] or null
  __nv_bfloat16* out;            // [M][ld_out]
  int ld_res, ld_out;
  int epilogue;
  int n_out_total;               // output columns (N, or N/2 for GEGLU)
  int tma_epilogue;              // 1: smem-staged TMA-store epilogue, 0: direct stores
  float* gn_partial;             // optional [ceil(M/32)][n_out_total][2] GroupNorm pre-reduction of the output
  // LayerNorm folded into the GEMMs around it:  LN(x) W^T + b = rstd_row * (x (gamma.W)^T - mu_row s + std_row b'),
  // s_n = sum_k gamma_k W_nk, b' = W beta + b.  The two rank-1 terms ride through the tensor core as ONE extra K chunk:
  // A gets a side tensor [M][64] holding (-mu, std) as bf16 hi/lo pairs (ln_side_kernel, norm.cu), B gets 64 extra
  // columns holding (s, b') the same way, so the epilogue only multiplies by rstd_row.
  float* ln_stats_out;           // producer: optional [M][2 * n_tiles][2] per-row (sum, sumsq) partials of the OUTPUT
  const float* row_scale;        // consumer: optional [M] per-row factor applied to the accumulator (no bias added)
  int gelu_tanh;                 // GEGLU with row_scale: 1 = one-MUFU tanh form of the GELU
  int pair;                      // 1: CTA-pair kernel (cta_group::2): work item = two adjacent M tiles x one N tile,
                                 //    CTA r of the pair owns M tile 2j + r and fetches B rows [r, r + 1) * block_n / 2
  int total_work;                // work items of the launch: m_tiles * n_tiles, or ceil(m_tiles / 2) * n_tiles
};

struct GemmOp {                  // host-side description; pointers are borrowed
  const void* a0 = nullptr; int c0 = 0, ld0 = 0;
  const void* a1 = nullptr; int c1 = 0, ld1 = 0;
  int n_img = 1, H = 1, W = 1;
  const void* w = nullptr; int N = 0, taps = 1;
  const float* bias = nullptr;
  const float* row_bias = nullptr;
  const void* residual = nullptr; int ld_res = 0;
  void* out = nullptr; int ld_out = 0;
  int epilogue = kEpiNone;
  int block_n = 0;               // 0 = choose
  int stride = 1;                // 2: 3x3 stride-2 pad-1 convolution; H, W are the OUTPUT extents, the input is 2H x 2W
  int upsample = 0;              // 1: 3x3 convolution of the nearest-2x upsampled input; H, W are the SOURCE extents,
                                 //    the output is 2H x 2W and w holds the 16 phase-combined 2x2 taps [16][N][K]
  float* gn_partial = nullptr;
  float* ln_stats_out = nullptr;           // see GemmParams
  const float* row_scale = nullptr;
  int pair = -1;                 // -1 = choose, 0 / 1 = force the one-CTA / CTA-pair kernel
};

struct GemmPlan {
  GemmParams p;
  int grid = 0;
  size_t smem = 0;
  double flops = 0;              // algorithmic 2*M*N*K
};

int gemm_choose_block_n(int N, int n_img, int H, int W, int epilogue);
int gemm_plan(const GemmOp& op, GemmPlan* plan);
int gemm_launch(const GemmPlan& plan, cudaStream_t stream);

}  // namespace sonic
