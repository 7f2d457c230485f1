// tcgen05 implicit-GEMM kernel: conv3x3 / conv1x1 / Linear for the SD UNet (see gemm.cuh).
//
// Warp roles (320 threads, one persistent CTA per SM):
//   warp 0      TMA producer   : A box (128 pixels x 64 ch, shifted per filter tap) + B box
//   warp 1      MMA issuer     : tcgen05.mma kind::f16, M=128, N=block_n, K=16 x4 per stage;
//                                owns the 512-column TMEM allocation (2 accumulator buffers).  The whole
//                                warp runs the loop (uniform registers), one elected lane issues.
//   warps 2..9  epilogue       : two warps per TMEM lane quarter; each handles 32 rows x 32-column chunks:
//                                tcgen05.ld -> bias / per-image bias / residual / GEGLU -> bf16 -> a private
//                                64B-swizzled staging buffer in shared memory -> TMA tile store.  The
//                                residual tile arrives the same way (TMA load into the staging buffer, L2
//                                prefetch one tile ahead), so global traffic is whole 64-byte row segments
//                                instead of one 16-byte piece per lane (a clock64 trace showed the old
//                                scattered-store epilogue taking 5800 cycles per 128x160 tile against a
//                                2600-cycle main loop at K=320).
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty pair (MMA <-> epilogue),
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include "gemm.cuh"

#include <algorithm>
#include <cstdlib>

namespace sonic {

namespace {

constexpr int kGemmThreads = 320;         // TMA warp + MMA warp + 8 epilogue warps
constexpr int kTileM = 128;
constexpr int kTileK = 64;                 // bf16 elements = one 128B swizzle row
constexpr int kABytes = kTileM * kTileK * 2;
constexpr int kMaxAcc = 4;                 // TMEM accumulator buffers: 512 columns / block_n rounded to 32 (2 .. 4); more
                                           // buffers let the MMA warp run further ahead of a bursty epilogue
constexpr int kMaxGranules = 8;            // 16-column granules per epilogue warp (block_n 256 / 2 / 16)
constexpr int kEpiWarps = 8;
constexpr int kChunkCols = 32;             // output columns per staged chunk (64 B of bf16: one swizzle-64B row)
constexpr int kStgBufBytes = 32 * kChunkCols * 2;       // 32 rows x 64 B
#ifndef SONIC_STG_BUFS
#define SONIC_STG_BUFS 4
#endif
#ifndef SONIC_STG_PENDING
#define SONIC_STG_PENDING 2
#endif
#ifndef SONIC_HINT_ACC
#define SONIC_HINT_ACC 128
#endif
#ifndef SONIC_HINT_FULL
#define SONIC_HINT_FULL 64
#endif
#ifndef SONIC_HINT_TF
#define SONIC_HINT_TF 64
#endif
constexpr uint32_t kHintAcc = SONIC_HINT_ACC, kHintFull = SONIC_HINT_FULL, kHintTf = SONIC_HINT_TF;   // suspend hints (ns): MMA warp's
                                           // waits for a free accumulator / a full stage, epilogue's wait for a full accumulator
constexpr int kStgBufs = SONIC_STG_BUFS;   // staging buffers per epilogue warp: a ring -- residual chunks are requested
constexpr int kStgPending = SONIC_STG_PENDING;   // kStgBufs - kStgPending chunks ahead (across tile boundaries) while up to
                                           // kStgPending tile stores drain behind
constexpr int kStgBytes = kEpiWarps * kStgBufs * kStgBufBytes;   // 64 KB (6 buffers / 96 KB cost two pipeline stages: slower)

// GELU (erf form, what diffusers' GEGLU computes) as x * sigmoid(2u), u = x (a + b x^2 + c x^4): the
// tanh-form GELU with its inner polynomial re-fitted (minimax, tools/fit_gelu.py) against the ERF form:
// max |error| 2.6e-5 over all x, an order of magnitude below bf16 resolution of the output.  The
// sigmoid form has no 1 - tanh cancellation for negative x.  9 instructions, 2 MUFU (ex2, rcp);
// -2 log2(e) is folded into the coefficients, |x| is clamped to 8 where the sigmoid is saturated.
__device__ __forceinline__ float gelu_erf(float x) {
  const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
  const float x2 = xc * xc;
  float v = fmaf(0.0010142650417074006f, x2, -0.1067757372109794f);
  v = fmaf(v, x2, -2.301121324206351f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * xc));
  return __fdividef(x, 1.0f + e);
}

// The same function with ONE MUFU op: x * sigmoid(2u) == h + h * tanh(u), h = x / 2, through tanh.approx.f32 -- 7
// instructions instead of 12.  tanh.approx has ~2^-11 relative error, which becomes up to ~1e-3 ABSOLUTE for
// strongly negative gates (|h| * error while the exact result is ~0), against 2.6e-5 for the two-MUFU form.
__device__ __forceinline__ float gelu_erf_tanh(float x) {
  const float x2 = fminf(x * x, 64.0f);                        // the fitted polynomial is only monotone up to |x| ~ 10
  float v = fmaf(-0.000351517477f, x2, 0.0370056506f);         // the polynomial above times -ln(2) / 2
  v = fmaf(v, x2, 0.797507879f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

template <bool kTanh>
__device__ __forceinline__ float gelu_sel(float x) { return kTanh ? gelu_erf_tanh(x) : gelu_erf(x); }

#ifdef SONIC_GEMM_TRACE
// Debug timeline (compile with -DSONIC_GEMM_TRACE): clock64 stamps of CTA 0's pipeline events per tile.
__device__ long long g_gemm_trace[4][64][4];
#define GEMM_TRACE(role, i, k) do { if (blockIdx.x == 0 && (i) < 64) g_gemm_trace[role][i][k] = clock64(); } while (0)
#else
#define GEMM_TRACE(role, i, k) do { } while (0)
#endif

// kPair: the CTA-pair form (cta_group::2, launched as clusters of two CTAs = the two SMs of a TPC).  A work item is two
// adjacent M tiles x one N tile; CTA r of the pair loads the A tile of M tile 2j + r and HALF of the B tile (rows
// [r, r + 1) * block_n / 2), the leader's MMA warp issues M = 256 instructions for both, every accumulator lands in
// the TMEM of the SM that owns its rows, and each CTA runs its own epilogue.  Operand traffic into an SM drops from
// 16 KB + block_n * 128 B to 16 KB + block_n * 64 B per K chunk (the one-CTA form needs ~70 B/clk/SM at block_n = 160
// or 256, which is what the L2 -> SM path delivers with all SMs pulling).  Barriers: TMA bytes of both CTAs are counted
// on the LEADER's full barrier; tcgen05.commit multicasts to the empty / tmem_full barriers of both CTAs; the
// epilogue warps of both CTAs arrive on the leader's tmem_empty barrier.
template <bool kPair>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_rows = kPair ? p.block_n / 2 : p.block_n;              // B rows this CTA fetches
  const int stage_bytes = kABytes + b_rows * kTileK * 2;             // multiple of 1024
  uint8_t* stg_base = smem + p.stages * stage_bytes;                 // epilogue staging, 32 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + kStgBytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + kMaxAcc;
  uint64_t* res_full = tmem_empty + kMaxAcc;                         // [kEpiWarps][kStgBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + kEpiWarps * kStgBufs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.total_work;
  const int rank = kPair ? static_cast<int>(blockIdx.x & 1) : 0;     // cluster rank: 0 = leader
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto m_tile_of = [&](int tile) { return kPair ? 2 * (tile / p.n_tiles) + rank : tile / p.n_tiles; };
  const int k_chunks = p.k_chunks0 + p.k_chunks1;
  const int k_iters = p.taps * k_chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kMaxAcc; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    for (int i = 0; i < kEpiWarps * kStgBufs; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();                 // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // everything above may overlap the previous kernel's tail (SONIC_PDL)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_a0);
      tma_prefetch_desc(&p.tm_b);
      if (p.k_chunks1) tma_prefetch_desc(&p.tm_a1);
      if (p.mode == kModeStride2)
        for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.tm_x[i]);
      int stage = 0;
      uint32_t phase = 0;
      int ti = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers, ++ti) {
        GEMM_TRACE(0, ti, 0);
        const int n_tile = tile % p.n_tiles;
        int m_tile = m_tile_of(tile);
        int phase_a = 0, phase_b = 0, b_tap0 = 0;
        if (p.mode == kModeUpsample) {                       // tiles are phase-major: (phase, source-pixel tile)
          const int ph = m_tile / p.m_tiles_src;
          m_tile -= ph * p.m_tiles_src;
          phase_a = ph >> 1; phase_b = ph & 1; b_tap0 = ph * 4;
        }
        const int w0 = (m_tile % p.tiles_w) * p.tile_w;
        const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.tile_h;
        const int i0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.tile_n;
        const int ncol0 = n_tile * p.block_n + rank * b_rows;
        for (int tap = 0; tap < p.taps; ++tap) {
          int dy = 0, dx = 0;
          const CUtensorMap* ma = &p.tm_a0;
          if (p.mode == kModeStride2) {
            // input row 2h + ky - 1: ky = 0 -> odd row of pair h - 1, ky = 1 -> even row of pair h, ky = 2 -> odd row
            // of pair h; the parity views make each of them an ordinary (zero-filled at -1) TMA box
            const int ky = tap / 3, kx = tap % 3;
            dy = ky == 0 ? -1 : 0; dx = kx == 0 ? -1 : 0;
            const int par = (ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0);
            ma = par == 0 ? &p.tm_a0 : &p.tm_x[par - 1];
          } else if (p.mode == kModeUpsample) {
            dy = (tap >> 1) - 1 + phase_a; dx = (tap & 1) - 1 + phase_b;
          } else if (p.taps == 9) {
            dy = tap / 3 - 1; dx = tap % 3 - 1;
          }
          for (int ch = 0; ch < k_chunks; ++ch) {
            mbar_wait<512>(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * stage_bytes;
            const CUtensorMap* mk = ch < p.k_chunks0 ? ma : &p.tm_a1;
            const int kc = (ch < p.k_chunks0 ? ch : ch - p.k_chunks0) * kTileK;
            if (kPair) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);     // both CTAs' bytes
              tma_load_4d_pair(sa, mk, &full_bar[stage], kc, w0 + dx, h0 + dy, i0);
              tma_load_3d_pair(sa + kABytes, &p.tm_b, &full_bar[stage], ch * kTileK, ncol0, b_tap0 + tap);
            } else {
              mbar_expect_tx(&full_bar[stage], stage_bytes);
              tma_load_4d(sa, mk, &full_bar[stage], kc, w0 + dx, h0 + dy, i0);
              tma_load_3d(sa + kABytes, &p.tm_b, &full_bar[stage], ch * kTileK, ncol0, b_tap0 + tap);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        GEMM_TRACE(0, ti, 1);
      }
      pdl_launch_dependents();                   // this CTA's last operand loads are in flight: the next kernel may set up
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform loop (descriptors stay in uniform registers), one elected lane issues.  In a CTA pair only the
    // leader CTA's warp runs it (the instructions drive both tensor cores).
    const bool leader = elect_one() && rank == 0;
    const uint64_t desc0 = make_sw128_desc(smem_u32(smem), 16, 1024);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    int ti = 0;
    for (int tile = worker; tile < total_tiles && rank == 0; tile += n_workers, ++ti) {
      if (lane == 0) GEMM_TRACE(1, ti, 0);
      mbar_wait<kHintAcc>(&tmem_empty[acc], acc_phase ^ 1);
      if (lane == 0) GEMM_TRACE(1, ti, 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
      for (int kit = 0; kit < k_iters; ++kit) {
        mbar_wait<kHintFull>(&full_bar[stage], phase);
        if (kit == 0 && lane == 0) GEMM_TRACE(1, ti, 2);
        tc_fence_after();
        const uint64_t da = desc0 + static_cast<uint64_t>((stage * stage_bytes) >> 4);
        const uint64_t db = da + (kABytes >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) {  // +32 B per K=16 step inside the 128B atom
            if (kPair) umma_bf16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kit | k) != 0);
            else umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kit | k) != 0);
          }
          if (kPair) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (leader) {
        if (kPair) umma_commit_pair(&tmem_full[acc]);
        else umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (lane == 0) GEMM_TRACE(1, ti, 3);
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
    }
  } else if (p.tma_epilogue) {
    // ------------------------------------------------------------------ epilogue, staged through smem
    const int ew = warp - 2;
    const int quarter = warp & 3;                // TMEM lane quarter this warp may read
    const int col_half0 = ew >> 2;               // even / odd 32-column chunks (swapped every other tile, below)
    const bool geglu = p.epilogue == kEpiGeglu;
    const bool has_res = p.residual != nullptr;
    const int out_cols = geglu ? p.block_n / 2 : p.block_n;
    const int n_chunks = out_cols / kChunkCols;
    uint8_t* stg = stg_base + ew * (kStgBufs * kStgBufBytes);
    uint64_t* r_full = res_full + ew * kStgBufs;
    const uint32_t sw = (lane >> 1) & 3;         // swizzle-64B: 16-byte piece j of row r sits at j ^ ((r >> 1) & 3)
    uint32_t slot = 0;                           // chunks staged so far: buffer = slot % kStgBufs, phase = (slot / kStgBufs) & 1
    int acc = 0;
    uint32_t acc_phase = 0;
    int ti = 0;
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_out);
      if (has_res) tma_prefetch_desc(&p.tm_res);
      if (p.mode == kModeUpsample)
        for (int i = 0; i < 3; ++i) tma_prefetch_desc(&p.tm_x[i]);
    }
    // Folded LayerNorm, consumer side: the mean / bias terms arrive THROUGH THE TENSOR CORE (an extra K chunk, see
    // gemm.cuh), so all that is left here is one multiply by this lane's row scale (rstd).  The scale of the NEXT tile
    // is requested a tile ahead: its L2 latency hides behind the current tile.
    float rs_next = 1.f;
    if (p.row_scale && worker < total_tiles)
      rs_next = __ldcg(p.row_scale + min(m_tile_of(worker) * kTileM + quarter * 32 + lane, p.M - 1));
    // Residual ring (lane 0 only): a load cursor runs kStgBufs - 1 chunks ahead of the chunk being processed, across
    // tile boundaries.  A clock64 trace of the K = 320 +residual projection showed the epilogue as the bottleneck
    // (5000 cycles per tile against 1850 of MMA): each chunk waited ~1000 cycles for its residual box and each tile
    // start ~1000 cycles for a store to release its buffer, because the small epilogue TMA operations queue behind
    // the operand boxes of the producer warp.  With three residual boxes in flight those waits overlap.
    int l_tile = worker, l_ti = 0, l_i = 0;
    uint32_t l_slot = 0;
    auto issue_next_residual = [&]() {           // lane 0
      while (l_tile < total_tiles) {
        const int half = (n_chunks & 1) ? (col_half0 ^ (l_ti & 1)) : col_half0;
        const int oc0 = (l_tile % p.n_tiles) * out_cols;
        const int c = half + 2 * l_i;
        if (c < n_chunks && oc0 + c * kChunkCols < p.n_out_total) {
          const uint32_t b = l_slot % kStgBufs;
          mbar_expect_tx(&r_full[b], kStgBufBytes);
          tma_load_2d(stg + b * kStgBufBytes, &p.tm_res, &r_full[b], oc0 + c * kChunkCols,
                      m_tile_of(l_tile) * kTileM + quarter * 32);
          ++l_i;
          ++l_slot;
          return;
        }
        l_tile += n_workers;
        ++l_ti;
        l_i = 0;
      }
    };
    if (has_res && lane == 0)
      for (int k = 0; k < kStgBufs - kStgPending; ++k) issue_next_residual();
    for (int tile = worker; tile < total_tiles; tile += n_workers, ++ti) {
      if (threadIdx.x == 64) GEMM_TRACE(2, ti, 0);
      const int n_tile = tile % p.n_tiles;
      int m_tile = m_tile_of(tile);
      int up_phase = 0;
      if (p.mode == kModeUpsample) {                         // phase-major tiles over the SOURCE pixels
        up_phase = m_tile / p.m_tiles_src;
        m_tile -= up_phase * p.m_tiles_src;
      }
      const int row0 = m_tile * kTileM + quarter * 32;       // tiles are 128 consecutive rows of [M][N]
      const int ncol0 = n_tile * p.block_n;                  // B-row / bias index base
      const int ocol0 = n_tile * out_cols;                   // output column base
      // With an odd chunk count (block_n = 160: 5 chunks) one warp of each quarter gets 3 chunks and the other 2.
      // The two swap roles on every tile, so over two tiles both do 5 -- the double-buffered accumulators let the
      // lighter warp run ahead into the next tile instead of idling at tmem_full (ncu: 26 % of epilogue time).
      const int col_half = (n_chunks & 1) ? (col_half0 ^ (ti & 1)) : col_half0;
      int my_n = 0;                                          // chunks col_half, col_half + 2, ... inside the matrix
      for (int c = col_half; c < n_chunks && ocol0 + c * kChunkCols < p.n_out_total; c += 2) ++my_n;
      const float rs = rs_next;
      if (p.row_scale) {
        const int tile2 = tile + n_workers;
        if (tile2 < total_tiles)
          rs_next = __ldcg(p.row_scale + min(m_tile_of(tile2) * kTileM + quarter * 32 + lane, p.M - 1));
      }
      float rs_sum = 0.f, rs_sq = 0.f;                       // producer side: statistics of this lane's output row
      if (threadIdx.x == 64) GEMM_TRACE(2, ti, 1);
      mbar_wait<kHintTf>(&tmem_full[acc], acc_phase);
      if (threadIdx.x == 64) GEMM_TRACE(2, ti, 2);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * p.acc_stride;
      if (my_n == 0) {
        __syncwarp();
        if (lane == 0) { if (kPair) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]); }
      }
      for (int i = 0; i < my_n; ++i) {
        const int c = col_half + 2 * i;
        const uint32_t b = slot % kStgBufs;
        uint8_t* buf = stg + b * kStgBufBytes;
        uint32_t v[32];
        float f[32];
        tmem_ld32(t_row + c * kChunkCols, v);
        if (i == 0 && threadIdx.x == 64) GEMM_TRACE(3, ti, 0);
        // p.bias is never null (gemm_plan substitutes a zero vector): one base pointer per chunk, constant offsets --
        // the per-load null test and 64-bit address arithmetic were ~17 % of the epilogue's issued instructions.
        const float4* bias_v = reinterpret_cast<const float4*>(p.bias + ncol0 + c * kChunkCols);
        if (geglu) {
          const float4* bias_g = reinterpret_cast<const float4*>(p.bias + ncol0 + out_cols + c * kChunkCols);
          uint32_t gt[32];
          tmem_ld32(t_row + out_cols + c * kChunkCols, gt);
          tmem_ld_wait();
          if (p.row_scale && p.gelu_tanh) {                    // bias already inside the accumulator
            // two outputs per instruction (packed fp32 pairs): this epilogue, not the MMA, bounds the 64x64 FFN projection
            const uint64_t rs2 = pack2(rs, rs);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint64_t x = fmul2(rs2, pack2(__uint_as_float(gt[j]), __uint_as_float(gt[j + 1])));
              float q0, q1;
              unpack2(fmul2(x, x), q0, q1);
              const uint64_t x2 = pack2(fminf(q0, 64.0f), fminf(q1, 64.0f));
              uint64_t u = ffma2(pack2(-0.000351517477f, -0.000351517477f), x2, pack2(0.0370056506f, 0.0370056506f));
              u = ffma2(u, x2, pack2(0.797507879f, 0.797507879f));
              float a0, a1, t0, t1;
              unpack2(fmul2(u, x), a0, a1);
              asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
              asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
              const uint64_t h = fmul2(x, pack2(0.5f, 0.5f));
              const uint64_t gl = ffma2(h, pack2(t0, t1), h);
              const uint64_t val = fmul2(rs2, pack2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
              unpack2(fmul2(val, gl), f[j], f[j + 1]);
            }
          } else if (p.row_scale) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              f[j] = (rs * __uint_as_float(v[j])) * gelu_erf(rs * __uint_as_float(gt[j]));
          } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(bias_v + (j >> 2)), bg = __ldg(bias_g + (j >> 2));
            f[j] = (__uint_as_float(v[j]) + bv.x) * gelu_erf(__uint_as_float(gt[j]) + bg.x);
            f[j + 1] = (__uint_as_float(v[j + 1]) + bv.y) * gelu_erf(__uint_as_float(gt[j + 1]) + bg.y);
            f[j + 2] = (__uint_as_float(v[j + 2]) + bv.z) * gelu_erf(__uint_as_float(gt[j + 2]) + bg.z);
            f[j + 3] = (__uint_as_float(v[j + 3]) + bv.w) * gelu_erf(__uint_as_float(gt[j + 3]) + bg.w);
          }
          }
        } else if (p.row_scale) {
          tmem_ld_wait();
          const uint64_t rs2 = pack2(rs, rs);
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            unpack2(fmul2(rs2, pack2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]))), f[j], f[j + 1]);
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {                    // two bias adds per instruction (packed fp32 pairs)
            const float4 bv = __ldg(bias_v + (j >> 2));
            unpack2(fadd2(pack2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), pack2(bv.x, bv.y)), f[j], f[j + 1]);
            unpack2(fadd2(pack2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), pack2(bv.z, bv.w)), f[j + 2], f[j + 3]);
          }
        }
        if (i == 0 && threadIdx.x == 64) GEMM_TRACE(3, ti, 1);
        if (p.epilogue == kEpiQuickGelu) {                   // x * sigmoid(1.702 x), CLIP's MLP activation
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f[j] * -2.4554669595930156f));   // -1.702 * log2(e)
            f[j] = __fdividef(f[j], 1.0f + e);
          }
        }
        if (i == my_n - 1) {                                 // every TMEM read of this tile is done
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (kPair) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]); }
        }
        if (p.row_bias) {
          const int img = min((row0 + lane) / (p.H * p.W), p.n_img - 1);
          const float* rb = p.row_bias + static_cast<size_t>(img) * p.N + ocol0 + c * kChunkCols;
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(rb + j);
        }
        uint8_t* my_row = buf + lane * 64;
        if (has_res) {
          mbar_wait<64>(&r_full[b], (slot / kStgBufs) & 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rr = *reinterpret_cast<const uint4*>(my_row + ((q ^ sw) << 4));
            const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              unpack2(fadd2(pack2(f[8 * q + 2 * e], f[8 * q + 2 * e + 1]), pack2(bf16_lo(rw[e]), bf16_hi(rw[e]))),
                      f[8 * q + 2 * e], f[8 * q + 2 * e + 1]);
          }
        } else {
          if (lane == 0) bulk_wait_read<kStgBufs - 1>();     // the store that last used this buffer has read it
          __syncwarp();
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(my_row + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        if (p.ln_stats_out) {                                // of the bf16-ROUNDED values: what the consumer multiplies
          uint64_t a2 = pack2(0.f, 0.f), q2 = pack2(0.f, 0.f);   // (even, odd) column accumulators, as before: a0 / a1, q0 / q1
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint64_t x = pack2(bf16_lo(pk[j]), bf16_hi(pk[j]));
            a2 = fadd2(a2, x);
            q2 = ffma2(x, x, q2);
          }
          float a0, a1, q0, q1;
          unpack2(a2, a0, a1);
          unpack2(q2, q0, q1);
          rs_sum += a0 + a1;
          rs_sq += q0 + q1;
        }
        if (i == 0 && threadIdx.x == 64) GEMM_TRACE(3, ti, 2);
        fence_proxy_async_smem();
        __syncwarp();
        if (p.gn_partial) {
          // GroupNorm pre-reduction: lane j sums column j of the staged (bf16-rounded) 32 x 32 chunk
          const int rmax = min(32, p.M - row0);
          const uint8_t* colp = buf + (lane & 7) * 2;
          const uint32_t q = lane >> 3;
          float sa[2] = {0.f, 0.f}, sq[2] = {0.f, 0.f};
#pragma unroll 8
          for (int r = 0; r < rmax; ++r) {
            const uint16_t h = *reinterpret_cast<const uint16_t*>(colp + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
            const float x = __uint_as_float(static_cast<uint32_t>(h) << 16);
            sa[r & 1] += x;
            sq[r & 1] = fmaf(x, x, sq[r & 1]);
          }
          const int col = ocol0 + c * kChunkCols + lane;
          // slot of this 32-row block among the per-image partials of the OUTPUT tensor (any order inside an image)
          size_t blk = static_cast<size_t>(row0 >> 5);
          if (p.mode == kModeUpsample) {
            const int hw = p.H * p.W, img = row0 / hw;
            blk = static_cast<size_t>(img) * (hw >> 3) + up_phase * (hw >> 5) + ((row0 - img * hw) >> 5);
          }
          if (col < p.n_out_total && rmax > 0)
            *reinterpret_cast<float2*>(p.gn_partial + (blk * p.n_out_total + col) * 2) =
                make_float2(sa[0] + sa[1], sq[0] + sq[1]);
        }
        if (lane == 0) {
          if (p.mode == kModeUpsample) {                     // 32 source pixels -> pixels (2h + a, 2w + b) of the output
            const int hw = p.H * p.W, img = row0 / hw, rem = row0 - img * hw;
            tma_store_4d(up_phase == 0 ? &p.tm_out : &p.tm_x[up_phase - 1], buf, ocol0 + c * kChunkCols, rem % p.W,
                         rem / p.W, img);
          } else {
            tma_store_2d(&p.tm_out, buf, ocol0 + c * kChunkCols, row0);
          }
          bulk_commit();
          if (has_res) {                                     // the ring's next box goes into the buffer chunk P - kStgPending
            bulk_wait_read<kStgPending>();                   // used, once its store has read it
            issue_next_residual();
          }
        }
        if (i == 0 && threadIdx.x == 64) GEMM_TRACE(3, ti, 3);
        ++slot;
      }
      if (p.ln_stats_out && row0 + lane < p.M)               // one slot per (row, n_tile, column half): no atomics
        reinterpret_cast<float2*>(p.ln_stats_out)[(static_cast<size_t>(row0 + lane) * p.n_tiles + n_tile) * 2 + col_half] =
            make_float2(rs_sum, rs_sq);
      if (threadIdx.x == 64) GEMM_TRACE(2, ti, 3);
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all<0>();
  } else {
    // ------------------------------------------------------------------ epilogue, direct stores
    // (block_n not a multiple of 32, e.g. the 4 -> 16 padded conv_out, or non-contiguous tiles)
    const int quarter = warp & 3;                // TMEM lane quarter this warp may read
    const int col_half = (warp - 2) >> 2;        // warps 2-5 take the low columns, 6-9 the high ones
    const int r = quarter * 32 + lane;           // row inside the 128-row tile
    const bool geglu = p.epilogue == kEpiGeglu;
    const int out_cols = geglu ? p.block_n / 2 : p.block_n;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < total_tiles; tile += n_workers) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = m_tile_of(tile);
      const int w = (m_tile % p.tiles_w) * p.tile_w + r % p.tile_w;
      const int h = ((m_tile / p.tiles_w) % p.tiles_h) * p.tile_h + (r / p.tile_w) % p.tile_h;
      const int img = (m_tile / (p.tiles_w * p.tiles_h)) * p.tile_n + r / (p.tile_w * p.tile_h);
      const bool row_ok = w < p.W && h < p.H && img < p.n_img;
      const size_t grow = (static_cast<size_t>(img) * p.H + h) * p.W + w;
      const int ncol0 = n_tile * p.block_n;                 // B-row / bias index base
      const int ocol0 = n_tile * out_cols;                  // output column base
      const int granules = out_cols / 16;
      const int g_begin = col_half == 0 ? 0 : (granules + 1) / 2;
      const int g_count = col_half == 0 ? (granules + 1) / 2 : granules - (granules + 1) / 2;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * p.acc_stride;
      for (int g = 0; g < g_count; ++g) {
        const int c = (g_begin + g) * 16;
        uint32_t v[16];
        tmem_ld16(t_row + c, v);
        float f[16];
        if (geglu) {
          uint32_t gt[16];
          tmem_ld16(t_row + out_cols + c, gt);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float bv = __ldg(p.bias + ncol0 + c + j);
            const float bg = __ldg(p.bias + ncol0 + out_cols + c + j);
            f[j] = (__uint_as_float(v[j]) + bv) * gelu_erf(__uint_as_float(gt[j]) + bg);
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + __ldg(p.bias + ncol0 + c + j);
        }
        const int oc = ocol0 + c;
        if (row_ok && oc < p.n_out_total) {
          if (p.row_bias) {
            const float* rb = p.row_bias + static_cast<size_t>(img) * p.N + oc;
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __ldg(rb + j);
          }
          if (p.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + grow * p.ld_res + oc);
            const uint4 r0 = __ldcg(rp), r1 = __ldcg(rp + 1);   // activations: coherent path (the kernel can overlap its producer's tail)
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[2 * j] += bf16_lo(rr[j]);
              f[2 * j + 1] += bf16_hi(rr[j]);
            }
          }
          uint4 o0, o1;
          o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);
          o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]);
          o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
          uint4* op = reinterpret_cast<uint4*>(p.out + grow * p.ld_out + oc);
          op[0] = o0;
          op[1] = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (kPair) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]); }
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while the peer still uses it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace
#ifdef SONIC_GEMM_TRACE
}  // namespace sonic
extern "C" int sonic_debug_gemm_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, sonic::g_gemm_trace, sizeof(long long) * 4 * 64 * 4) == cudaSuccess ? 0 : -1;
}
namespace sonic {
#endif
namespace {
int g_num_sms = 0;
bool g_attr_set = false;
constexpr int kZeroBiasLen = 16384;                // >= the widest layer (GEGLU 10240), padded tiles included
__device__ float g_zero_bias_dev[kZeroBiasLen];    // zero-initialised: the "bias" of bias-free layers
const float* g_zero_bias = nullptr;

int pair_enabled() {
  static const int on = [] { const char* e = getenv("SONIC_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  return on;
}

int pick_block_n(int N, int m_tiles, bool geglu, int num_sms, bool pair) {
  // Prefer wide tiles (A re-use, fewer smem bytes per MMA cycle) but avoid tail waves.
  static const int cands[] = {256, 224, 192, 160, 128, 96, 64, 32, 16};   // multiples of 32: staged epilogue chunks
  static const int model = [] { const char* e = getenv("SONIC_BN_MODEL"); return e ? atoi(e) : 1; }();
  int best = 0;
  double best_cost = 1e30;
  for (int bn : cands) {
    if (bn > N && bn != 16) continue;
    if (bn == 16 && N >= 32) continue;            // 16 only for the padded 4 -> 16 conv_out (direct-store epilogue)
    if (geglu && (bn % 64 != 0 || N % bn != 0)) continue;
    const int n_tiles = (N + bn - 1) / bn;
    double cost;
    if (pair && model) {
      // CTA pairs: work items of two M tiles on num_sms / 2 clusters; an MMA of width bn costs bn / 2 + ~48 cycles
      // (round-1 microbenchmark), so wide tiles are worth a partly filled last wave more often than "bn + 12" says
      const long items = static_cast<long>((m_tiles + 1) / 2) * n_tiles;
      const long workers = std::max(1, num_sms / 2);
      const long waves = (items + workers - 1) / workers;
      static const double ovh = [] { const char* e = getenv("SONIC_BN_OVH"); return e ? atof(e) : 51.0; }();
      cost = waves * (bn * 0.5 + ovh);
    } else {
      const long tiles = static_cast<long>(m_tiles) * n_tiles;
      const long waves = (tiles + num_sms - 1) / num_sms;
      // per-tile time ~ max(MMA cycles ~ bn, smem-feed cycles ~ (128+bn)/2) + fixed overhead
      const double per_tile = std::max<double>(bn, (128.0 + bn) * 0.5 * 1.1) + 12.0;
      cost = waves * per_tile;
    }
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best ? best : 16;
}

}  // namespace

int gemm_choose_block_n(int N, int n_img, int H, int W, int epilogue) {
  if (!g_num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_num_sms = 148;
  }
  int m_tiles;
  if (W >= kTileM || (H == 1 && n_img == 1)) {
    m_tiles = ((W + kTileM - 1) / kTileM) * H * n_img;
  } else {
    const int th = std::min(H, kTileM / W);
    const int tn = kTileM / (W * th);
    m_tiles = (H / th) * ((n_img + tn - 1) / tn);
  }
  return pick_block_n(N, m_tiles, epilogue == kEpiGeglu, g_num_sms, pair_enabled() && m_tiles >= 2 && g_num_sms % 2 == 0);
}

int gemm_plan(const GemmOp& op, GemmPlan* plan) {
  SONIC_REQUIRE(op.a0 && op.w && op.out, "gemm: null operand");
  SONIC_REQUIRE(op.taps == 1 || op.taps == 9, "gemm: taps must be 1 or 9 (got %d)", op.taps);
  SONIC_REQUIRE(op.stride == 1 || op.stride == 2, "gemm: stride must be 1 or 2");
  SONIC_REQUIRE(!(op.stride == 2 && op.upsample), "gemm: stride 2 and upsample are exclusive");
  SONIC_REQUIRE((op.stride == 1 && !op.upsample) || (op.taps == 9 && op.a1 == nullptr && op.residual == nullptr &&
                                                      op.row_bias == nullptr && op.epilogue == kEpiNone),
                "gemm: stride-2 / upsample convolutions are plain 3x3 convolutions of one tensor");
  const int mode = op.stride == 2 ? kModeStride2 : op.upsample ? kModeUpsample : kModeNormal;
  SONIC_REQUIRE(op.N % 16 == 0, "gemm: N=%d must be a multiple of 16", op.N);
  SONIC_REQUIRE(op.c0 % 8 == 0 && op.c1 % 8 == 0 && op.ld0 % 8 == 0 && op.ld1 % 8 == 0,
                "gemm: channel counts / strides must be multiples of 8");
  SONIC_REQUIRE(op.c0 % 64 == 0 || op.a1 == nullptr, "gemm: concat needs c0 %% 64 == 0");
  SONIC_REQUIRE(op.ld_out % 8 == 0 && (op.residual == nullptr || op.ld_res % 8 == 0),
                "gemm: output / residual stride must be a multiple of 8");
  if (!g_num_sms) {
    int dev = 0;
    SONIC_CUDA(cudaGetDevice(&dev));
    SONIC_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  GemmParams& p = plan->p;
  p = GemmParams{};
  const int K = op.c0 + (op.a1 ? op.c1 : 0);
  p.M = op.n_img * op.H * op.W;
  p.N = op.N;
  p.k_chunks0 = (op.c0 + kTileK - 1) / kTileK;
  p.k_chunks1 = op.a1 ? (op.c1 + kTileK - 1) / kTileK : 0;
  p.taps = mode == kModeUpsample ? 4 : op.taps;
  p.mode = mode;
  p.H = op.H; p.W = op.W; p.n_img = op.n_img;
  if (op.W >= kTileM || (op.H == 1 && op.n_img == 1)) {
    // rows of an image row (or of a plain [M][K] matrix, however short: the part of the 128-row box beyond the
    // tensor is zero-filled by the TMA unit and its output rows are clipped by the tile store)
    p.tile_w = kTileM; p.tile_h = 1; p.tile_n = 1;
  } else {
    SONIC_REQUIRE(kTileM % op.W == 0, "gemm: W=%d must divide 128", op.W);
    p.tile_w = op.W;
    p.tile_h = std::min(op.H, kTileM / p.tile_w);
    SONIC_REQUIRE((kTileM / p.tile_w) % p.tile_h == 0 && op.H % p.tile_h == 0,
                  "gemm: H=%d does not tile", op.H);
    p.tile_n = kTileM / (p.tile_w * p.tile_h);
  }
  p.tiles_w = (op.W + p.tile_w - 1) / p.tile_w;
  p.tiles_h = (op.H + p.tile_h - 1) / p.tile_h;
  p.tiles_img = (op.n_img + p.tile_n - 1) / p.tile_n;
  p.m_tiles_src = p.tiles_w * p.tiles_h * p.tiles_img;
  p.m_tiles = p.m_tiles_src * (mode == kModeUpsample ? 4 : 1);
  const bool geglu = op.epilogue == kEpiGeglu;
  const bool pair_shape = g_num_sms % 2 == 0 && p.m_tiles >= 2 && (mode != kModeUpsample || p.m_tiles_src % 2 == 0);
  p.block_n = op.block_n ? op.block_n
                         : pick_block_n(op.N, p.m_tiles, geglu, g_num_sms, pair_shape && op.pair != 0 && (op.pair > 0 || pair_enabled()));
  SONIC_REQUIRE(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "gemm: bad block_n %d",
                p.block_n);
  SONIC_REQUIRE(!geglu || (p.block_n % 32 == 0 && op.N % p.block_n == 0),
                "gemm: GEGLU needs block_n %% 32 == 0 and N %% block_n == 0");
  p.n_out_total = geglu ? op.N / 2 : op.N;
  // staged (TMA-store) epilogue: 32-column chunks of tiles that are 128 consecutive rows of [M][N]
  const int out_cols = geglu ? p.block_n / 2 : p.block_n;
  const bool flat = op.W < kTileM || op.W % kTileM == 0 || (op.H == 1 && op.n_img == 1);
  p.tma_epilogue = (out_cols % kChunkCols == 0 && flat) ? 1 : 0;
  p.gn_partial = op.gn_partial;
  p.ln_stats_out = op.ln_stats_out;
  p.row_scale = op.row_scale;
  {
    const char* e = getenv("SONIC_GELU_TANH");               // experiment switch: one-MUFU GELU in the folded GEGLU epilogue
    p.gelu_tanh = (e && e[0] == '0') ? 0 : 1;
  }
  SONIC_REQUIRE((op.ln_stats_out == nullptr && op.row_scale == nullptr) || p.tma_epilogue,
                "gemm: folded LayerNorm needs the staged epilogue (block_n %% 32 == 0, contiguous 128-row tiles)");
  SONIC_REQUIRE(op.row_scale == nullptr || (op.bias == nullptr && op.residual == nullptr && op.row_bias == nullptr),
                "gemm: row_scale (folded LayerNorm) carries its bias inside the product: no bias / residual / row_bias");
  SONIC_REQUIRE(op.epilogue != kEpiQuickGelu || p.tma_epilogue, "gemm: QuickGELU needs the staged epilogue");
  SONIC_REQUIRE(op.gn_partial == nullptr || p.tma_epilogue,
                "gemm: gn_partial needs the staged epilogue (block_n %% 32 == 0, contiguous 128-row tiles)");
  p.n_tiles = (op.N + p.block_n - 1) / p.block_n;
  // CTA pairs: two adjacent M tiles share one B tile.  Needs an even SM count, a B half of whole 8-row swizzle atoms,
  // phase-aligned tile pairs in the upsample form, and enough work that halving the worker count costs no wave.
  const int pair_env = pair_enabled();
  static const int pair_min = [] { const char* e = getenv("SONIC_GEMM_PAIR_MIN"); return e ? atoi(e) : 0; }();
  const bool pair_ok = p.block_n % 32 == 0 && g_num_sms % 2 == 0 && p.m_tiles >= 2 &&
                       (mode != kModeUpsample || p.m_tiles_src % 2 == 0);
  if (op.pair >= 0) {
    SONIC_REQUIRE(!op.pair || pair_ok, "gemm: the CTA-pair kernel cannot run this shape");
    p.pair = op.pair;
  } else {
    p.pair = pair_env && pair_ok && static_cast<long>(p.m_tiles) * p.n_tiles >= static_cast<long>(pair_min) * g_num_sms;
  }
  p.total_work = p.pair ? ((p.m_tiles + 1) / 2) * p.n_tiles : p.m_tiles * p.n_tiles;
  const int b_rows = p.pair ? p.block_n / 2 : p.block_n;
  const int stage_bytes = kABytes + b_rows * kTileK * 2;
  p.stages = std::max(2, std::min(8, (224 * 1024 - kStgBytes) / stage_bytes));
  p.idesc = make_idesc_bf16(p.pair ? 2 * kTileM : kTileM, p.block_n, false);
  p.acc_stride = (p.block_n + 31) / 32 * 32;
  p.n_acc = std::max(2, std::min(kMaxAcc, 512 / p.acc_stride));
  if (p.n_acc == 2) p.acc_stride = 256;
  if (!op.bias) {
    SONIC_REQUIRE(op.N <= kZeroBiasLen, "gemm: N=%d exceeds the zero-bias vector", op.N);
    if (!g_zero_bias) {
      void* zp = nullptr;
      SONIC_CUDA(cudaGetSymbolAddress(&zp, g_zero_bias_dev));
      g_zero_bias = static_cast<const float*>(zp);
    }
  }
  p.bias = op.bias ? op.bias : g_zero_bias;
  p.row_bias = op.row_bias;
  p.residual = static_cast<const __nv_bfloat16*>(op.residual);
  p.out = static_cast<__nv_bfloat16*>(op.out);
  p.ld_res = op.ld_res;
  p.ld_out = op.ld_out;
  p.epilogue = op.epilogue;

  {
    uint64_t dims[4] = {static_cast<uint64_t>(op.c0), static_cast<uint64_t>(op.W),
                        static_cast<uint64_t>(op.H), static_cast<uint64_t>(op.n_img)};
    uint64_t str[3] = {static_cast<uint64_t>(op.ld0) * 2, static_cast<uint64_t>(op.W) * op.ld0 * 2,
                       static_cast<uint64_t>(op.H) * op.W * op.ld0 * 2};
    uint32_t box[4] = {kTileK, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h),
                       static_cast<uint32_t>(p.tile_n)};
    if (mode == kModeStride2) {
      // four parity views of the 2H x 2W input: view (a, b) element (n, h, w) = input pixel (2h + a, 2w + b)
      const uint64_t ld = static_cast<uint64_t>(op.ld0), Wi = 2ull * op.W, Hi = 2ull * op.H;
      uint64_t st2[3] = {2 * ld * 2, 2 * Wi * ld * 2, Hi * Wi * ld * 2};
      for (int par = 0; par < 4; ++par) {
        const char* base = static_cast<const char*>(op.a0) + ((par >> 1) * Wi + (par & 1)) * ld * 2;
        if (int rc = encode_tensor_map(par == 0 ? &p.tm_a0 : &p.tm_x[par - 1], base, 4, dims, st2, box, 128)) return rc;
      }
    } else if (int rc = encode_tensor_map(&p.tm_a0, op.a0, 4, dims, str, box, 128)) {
      return rc;
    }
    if (op.a1) {
      dims[0] = op.c1;
      str[0] = static_cast<uint64_t>(op.ld1) * 2;
      str[1] = static_cast<uint64_t>(op.W) * op.ld1 * 2;
      str[2] = static_cast<uint64_t>(op.H) * op.W * op.ld1 * 2;
      if (int rc = encode_tensor_map(&p.tm_a1, op.a1, 4, dims, str, box, 128)) return rc;
    } else {
      p.tm_a1 = p.tm_a0;
    }
  }
  {
    const uint64_t Kp = static_cast<uint64_t>(p.k_chunks0 + p.k_chunks1) * kTileK;
    SONIC_REQUIRE(Kp == static_cast<uint64_t>(K) || op.a1 == nullptr,
                  "gemm: concat operands must be multiples of 64 channels");
    uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(op.N),
                        static_cast<uint64_t>(mode == kModeUpsample ? 16 : op.taps)};
    uint64_t str[2] = {static_cast<uint64_t>(K) * 2, static_cast<uint64_t>(op.N) * K * 2};
    uint32_t box[3] = {kTileK, static_cast<uint32_t>(b_rows), 1};
    if (int rc = encode_tensor_map(&p.tm_b, op.w, 3, dims, str, box, 128)) return rc;
  }
  p.tm_out = p.tm_b;
  p.tm_res = p.tm_b;
  if (mode == kModeUpsample) {
    // four phase views of the 2H x 2W output: view (a, b) element (n, h, w) = output pixel (2h + a, 2w + b); a staged
    // chunk of 32 consecutive source pixels is a (32 / tw) x tw patch of the view
    SONIC_REQUIRE(p.tma_epilogue, "gemm: the upsample convolution needs the staged epilogue (N %% 32 == 0)");
    const int tw = std::min(32, op.W), th = 32 / tw;
    SONIC_REQUIRE(32 % tw == 0 && op.W % tw == 0 && (op.H * op.W) % 32 == 0 && op.H % th == 0,
                  "gemm: upsample convolution needs 32-pixel chunks that tile the %dx%d source", op.H, op.W);
    const uint64_t ld = static_cast<uint64_t>(op.ld_out), Wo = 2ull * op.W, Ho = 2ull * op.H;
    uint64_t dims[4] = {static_cast<uint64_t>(p.n_out_total), static_cast<uint64_t>(op.W), static_cast<uint64_t>(op.H),
                        static_cast<uint64_t>(op.n_img)};
    uint64_t str[3] = {2 * ld * 2, 2 * Wo * ld * 2, Ho * Wo * ld * 2};
    uint32_t box[4] = {kChunkCols, static_cast<uint32_t>(tw), static_cast<uint32_t>(th), 1};
    for (int ph = 0; ph < 4; ++ph) {
      char* base = static_cast<char*>(op.out) + ((ph >> 1) * Wo + (ph & 1)) * ld * 2;
      if (int rc = encode_tensor_map(ph == 0 ? &p.tm_out : &p.tm_x[ph - 1], base, 4, dims, str, box, 64)) return rc;
    }
  } else if (p.tma_epilogue) {
    uint64_t dims[2] = {static_cast<uint64_t>(p.n_out_total), static_cast<uint64_t>(p.M)};
    uint64_t str[1] = {static_cast<uint64_t>(op.ld_out) * 2};
    uint32_t box[2] = {kChunkCols, 32};
    if (int rc = encode_tensor_map(&p.tm_out, op.out, 2, dims, str, box, 64)) return rc;
    if (op.residual) {
      str[0] = static_cast<uint64_t>(op.ld_res) * 2;
      if (int rc = encode_tensor_map(&p.tm_res, op.residual, 2, dims, str, box, 64)) return rc;
    }
  }
  plan->grid = p.pair ? 2 * std::min(p.total_work, g_num_sms / 2) : std::min(p.total_work, g_num_sms);
  plan->smem = static_cast<size_t>(p.stages) * stage_bytes + kStgBytes + 1024 /*align*/ + 1024 /*barriers*/;
  // algorithmic work of the operator (a 9-tap convolution of every OUTPUT pixel), whatever the kernel executes: the
  // phase form of the upsample convolution runs 4/9 of it
  // (likewise the side-tensor K chunk of a folded LayerNorm is not algorithmic work)
  plan->flops = 2.0 * p.M * (mode == kModeUpsample ? 4.0 : 1.0) * static_cast<double>(op.N) *
                (op.row_scale ? op.c0 : K) * op.taps;
  return 0;
}

int gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
  if (!g_attr_set) {
    SONIC_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SONIC_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_set = true;
  }
  if (plan.p.pair) {
    SONIC_CUDA(launch_kernel_ex(conv_gemm_kernel<true>, dim3(plan.grid), dim3(kGemmThreads), plan.smem, stream, 2, plan.p));
  } else if (pdl_enabled()) {
    SONIC_CUDA(launch_kernel_ex(conv_gemm_kernel<false>, dim3(plan.grid), dim3(kGemmThreads), plan.smem, stream, 1, plan.p));
  } else {
    conv_gemm_kernel<false><<<plan.grid, kGemmThreads, plan.smem, stream>>>(plan.p);
  }
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sonic
