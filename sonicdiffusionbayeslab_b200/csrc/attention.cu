// Flash attention on tcgen05 tensor cores for the UNet's self- and cross-attention layers.
//   O[b, s, h, :] = softmax(scale * Q[b, s, h, :] K[b, :, h, :]^T) V[b, :, h, :]
// Q/K/V are read in place from the projection GEMM outputs ([B*S, ld] rows, head h at column
// h*D) through strided 4-D TMA tensor maps; head dims that are not a multiple of 64 (40, 80,
// 160) are zero-filled by the TMA unit, never padded in HBM.
//
// Two kernels share the scheme below.  attention_kernel: one CTA = one 128-row query tile of one (batch, head)
// (head dim 80, at most 128 queries, and -- in loop mode -- the cross-attention layers).  attention2_kernel (further
// down): one CTA owns TWO query tiles and ping-pongs between them; two CTAs per SM for head dims <= 64 (the 64x64 UNet
// level and the CLIP towers), one CTA per SM with 256 TMEM columns per tile for head dims >= 96 (the 16x16 level).
// Keys are consumed in sub-tiles of 64 and everything between the two GEMMs lives in tensor memory:
//   TMEM columns  [0, 64)        S = Q K^T scores of a sub-tile, fp32; once a thread has read its row it
//                                overwrites columns [0, 32) with P = exp2(...) as bf16 pairs -- the A operand
//                                of the PV product, read by the tensor core straight from TMEM (".ts" MMA)
//                 [64, 64+kDPV)  O accumulator, fp32, accumulated across all keys by the MMA
// With 112 columns (d <= 64) a CTA allocates 128, so FOUR CTAs are resident per SM (4 x 48 KB of shared
// memory): the kernel is bound by the MUFU ex2 rate (16/clk/SM, 32*8*4096^2 exponentials per 64x64 layer)
// and a CTA's dependency chain  S(t) -> softmax(t) -> PV(t), S(t+1)  leaves the pipe idle for the MMA
// round trip; four independent CTAs per sub-partition keep it fed (clock64 traces: one softmax warp per
// sub-partition reaches 73% of the MUFU rate, two 86%).
// Warp roles (192 threads):
//   warps 0-3  softmax      : one query row per thread (TMEM lane == row, so row max / sum need no
//                             shuffles): load the row's 64 scores, max, exponentiate against a LAZY running
//                             reference m that is only raised when a row exceeds it by more than 2^8 (only
//                             then is O rescaled in TMEM with tcgen05.ld/st), write P.
//   warp 4     TMA producer : Q once, then K and V tiles of 64 keys (2-stage rings)
//   warp 5     MMA issuer   : O += P(t) V(t) (TS), then S(t+1) = Q K^T (SS); the tensor pipe executes in
//                             issue order, so S(t+1) overwrites the P(t) columns only after PV(t) has read
//                             them, and "S(t) complete" implies "PV(t-1) complete".
#include "ops.cuh"

#include <cstdlib>
#include <type_traits>

namespace sonic {

struct AttentionPlan {
  CUtensorMap tm_q, tm_k, tm_v;
  AttentionOp op;
  int dpv = 0;         // head dim rounded up to a supported MMA N of the PV product
  int atoms = 0;       // 64-column smem atoms per row (1, 2, 3)
  int qt = 1;          // 128-query tiles per CTA: 2 = the ping-pong kernel (head dim <= 64)
  bool resident = false;   // two-tile kernel with K / V resident in shared memory, one CTA per (batch, head)
  int tiles_per_cta = 1;   // one-tile kernel in loop mode (K / V resident, <= 2 key sub-tiles): query tiles per CTA
  size_t smem = 0;
  dim3 grid;
};

namespace {

constexpr int kAttThreads = 192;
// Softmax warps are 0-3 (warp w owns TMEM lanes 32w..32w+31); the two service warps get the HIGHEST warp
// ids because the sub-partition arbiter prefers high ids: their rare instructions (TMA issue, MMA issue --
// the critical path between two softmax sub-tiles) must not queue behind the math warps.
constexpr int kWarpTma = 4;
constexpr int kWarpMma = 5;
constexpr int kBlockQ = 128;
constexpr int kQAtomBytes = kBlockQ * 128;      // 128 rows x 64 bf16
constexpr int kSub = 64;                        // keys per sub-tile (TMA tile == S tile == P tile)
constexpr int kKvAtomBytes = kSub * 128;        // 64 rows x 64 bf16
constexpr float kLazyLog2 = 8.0f;               // P <= 2^8 before the reference is raised

struct AttParams {
  CUtensorMap tm_q, tm_k, tm_v;
  __nv_bfloat16* o;
  int ld_o, seq_q, seq_k, head_dim, atoms, causal;
  float scale_log2;
  uint32_t idesc_s, idesc_pv;
  int tail_w;              // keys of the last sub-tile rounded up to 16 / 32 / 64: its S, softmax and PV only span these
  uint32_t idesc_s_tail;   // S product of the last sub-tile (N = tail_w)
  int tiles_per_cta;       // attention_kernel, K / V resident (<= 2 key sub-tiles): query tiles one CTA walks over
  int speculate;           // exponentiate full sub-tiles after the first without the row-maximum pass (see softmax_sub)
};

#ifdef SONIC_ATT_TRACE
// Debug timeline (compile with -DSONIC_ATT_TRACE): clock64 stamps of one CTA's pipeline events.
__device__ long long g_att_trace[8][160][4];
#define ATT_TRACE(w, t, k) do { if (blockIdx.x == 5 && blockIdx.y == 3 && blockIdx.z == 7 && (t) < 160 && lane == 0) \
    g_att_trace[w][t][k] = clock64(); } while (0)
#else
#define ATT_TRACE(w, t, k) do { } while (0)
#endif

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- the exponentials of a score row, two at a time ----------------------------------------------------------
// The d <= 64 kernels are bound by the MUFU unit (16 ex2 per clock and SM) AND short of issue slots (stall sampling:
// 17 % selected, 29 % fixed-latency waits with two softmax warps per sub-partition), so the arithmetic around the
// exponentials is done on register PAIRS with Blackwell's packed fp32 instructions (fma.rn.f32x2 / add.rn.f32x2 ->
// SASS FFMA2 / FADD2): one instruction scales two scores, one adds two exponentials to the row sum -- 5 issue slots
// per pair instead of 7 -- and kPolyOf8 of every eight pairs take their exponentials from the FMA pipe instead of
// the MUFU unit: round-to-nearest split x = j + r (magic-number add), degree-3 minimax polynomial of 2^r on
// [-1/2, 1/2] (relative error 7.5e-5, a 26th of the bf16 rounding P gets anyway), j added to the exponent field.
// The clamp keeps the exponent arithmetic in range: the low side flushes to 2^-125, the high side stays far above
// the speculative pass's 2^8 bound, so an out-of-range score still forces the checked path.  The speculative and
// the checked pass share this code, so they stay bit-identical.
#ifndef SONIC_ATT_POLY
#define SONIC_ATT_POLY 1
#endif
constexpr int kPolyOf8 = SONIC_ATT_POLY;
__device__ __forceinline__ constexpr bool poly_pair(int j) {          // pair j of a row (compile-time after unrolling)
  return kPolyOf8 >= 8 ? true
       : kPolyOf8 == 4 ? (j & 1) == 1
       : kPolyOf8 == 3 ? ((j & 7) == 2 || (j & 7) == 5 || (j & 7) == 7)
       : kPolyOf8 == 2 ? (j & 3) == 3
       : kPolyOf8 == 1 ? (j & 7) == 7
       : false;
}
__device__ __forceinline__ void poly_exp2_pair(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fminf(fmaxf(x0, -125.0f), 125.0f);
  x1 = fminf(fmaxf(x1, -125.0f), 125.0f);
  const uint64_t xc = pack2(x0, x1);
  const uint64_t t2 = fadd2(xc, pack2(12582912.0f, 12582912.0f));    // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const uint64_t j2 = fadd2(t2, pack2(-12582912.0f, -12582912.0f));
  const uint64_t r2 = ffma2(j2, pack2(-1.0f, -1.0f), xc);
  uint64_t p2 = ffma2(pack2(0x1.c3f6a6p-5f, 0x1.c3f6a6p-5f), r2, pack2(0x1.f0ddccp-3f, 0x1.f0ddccp-3f));
  p2 = ffma2(p2, r2, pack2(0x1.62f31ap-1f, 0x1.62f31ap-1f));
  p2 = ffma2(p2, r2, pack2(0x1.fff694p-1f, 0x1.fff694p-1f));
  float p0, p1, t0, t1;
  unpack2(p2, p0, p1);
  unpack2(t2, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// kN scores v[0 .. kN) (row columns col0 ...) -> bf16 pairs pk[0 .. kN / 2), exponentials added to the four pair
// accumulators.  scale2 = (c, c), nms2 = (-m c, -m c).
template <int kN, bool kMask>
__device__ __forceinline__ void exp_row(const uint32_t* v, uint32_t* pk, int col0, int valid, uint64_t scale2, uint64_t nms2,
                                        uint64_t (&acc)[4]) {
#pragma unroll
  for (int i = 0; i < kN; i += 2) {
    const uint64_t x2 = ffma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), scale2, nms2);
    float e0, e1;
    if (poly_pair(i >> 1)) {
      poly_exp2_pair(x2, e0, e1);
    } else {
      float x0, x1;
      unpack2(x2, x0, x1);
      e0 = fast_exp2(x0);
      e1 = fast_exp2(x1);
    }
    if (kMask && col0 + i >= valid) e0 = 0.f;
    if (kMask && col0 + i + 1 >= valid) e1 = 0.f;
    acc[(i >> 1) & 3] = fadd2(acc[(i >> 1) & 3], pack2(e0, e1));
    pk[i >> 1] = pack_bf16(e0, e1);
  }
}
__device__ __forceinline__ float sum_acc(const uint64_t (&acc)[4]) {
  float a[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) unpack2(acc[k], a[2 * k], a[2 * k + 1]);
  return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

template <int kDPV>
__global__ void __launch_bounds__(kAttThreads, (kDPV <= 64 ? 4 : 2))
attention_kernel(const __grid_constant__ AttParams p) {
  constexpr int kTmemCols = kSub + kDPV <= 128 ? 128 : 256;
  static_assert(kSub + kDPV <= 256, "S/P + O must fit the CTA's TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int q_bytes = p.atoms * kQAtomBytes;
  const int kv_bytes = p.atoms * kKvAtomBytes;
  uint8_t* sm_q = smem;
  uint8_t* sm_k = sm_q + q_bytes;                // 2 stages
  uint8_t* sm_v = sm_k + 2 * kv_bytes;           // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_v + 2 * kv_bytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // S(t) complete in TMEM (and every earlier MMA, incl. PV(t-1))
  uint64_t* p_full = bars + 10;       // P(t) written over S(t), O rescaled if it had to be
  uint64_t* o_done = bars + 11;       // last PV of the current query tile complete
  uint64_t* q_empty = bars + 12;      // every S product of the current query tile complete: Q may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_sub = (p.seq_k + kSub - 1) / kSub;
  const int k_steps_s = (p.head_dim + 15) / 16;
  // Loop mode (short key sequences, i.e. the cross-attention layers): K and V of this (batch, head) stay in the two
  // ring stages for the whole CTA, which walks over `tiles_per_cta` query tiles -- the per-CTA costs (TMEM
  // allocation, barrier set-up, first K / V loads) are paid once, and with four CTAs per SM sixteen softmax warps
  // hide each other's S -> softmax -> PV round trips.  tiles_per_cta == 1 is the plain one-tile kernel.
  const bool resident = p.tiles_per_cta > 1;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  const int total_tiles = (p.seq_q + kBlockQ - 1) / kBlockQ;
  const int n_tiles = min(p.tiles_per_cta, total_tiles - tile0);
  // a ragged last sub-tile (77 keys: 64 + 13) only spans 32 score columns in the softmax and the PV product
  const int w_last = p.tail_w < kSub ? 32 : kSub;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_done, 1); mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;
  const uint32_t tmem_o = tmem_s + kSub;
  pdl_wait();                                    // the set-up above may overlap the previous kernel's tail (SONIC_PDL)

  if (warp == kWarpTma) {
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v);
      for (int i = 0; i < n_tiles; ++i) {
        if (i > 0) mbar_wait<256>(q_empty, (i - 1) & 1);
        mbar_expect_tx(q_full, q_bytes);
        for (int a = 0; a < p.atoms; ++a)
          tma_load_4d(sm_q + a * kQAtomBytes, &p.tm_q, q_full, a * 64, head, (tile0 + i) * kBlockQ, batch);
        if (i > 0) continue;                     // K / V: once (resident) -- or the streaming ring of the one-tile form
        for (int t = 0; t < n_sub; ++t) {
          const int st = t & 1;
          const uint32_t ph = ((t >> 1) & 1) ^ 1;
          if (!resident) mbar_wait<512>(&k_empty[st], ph);
          mbar_expect_tx(&k_full[st], kv_bytes);
          for (int a = 0; a < p.atoms; ++a)
            tma_load_4d(sm_k + st * kv_bytes + a * kKvAtomBytes, &p.tm_k, &k_full[st], a * 64, head, t * kSub, batch);
          if (!resident) mbar_wait<512>(&v_empty[st], ph);
          mbar_expect_tx(&v_full[st], kv_bytes);
          for (int a = 0; a < p.atoms; ++a)
            tma_load_4d(sm_v + st * kv_bytes + a * kKvAtomBytes, &p.tm_v, &v_full[st], a * 64, head, t * kSub, batch);
        }
      }
      pdl_launch_dependents();                   // every load of this CTA has been issued
    }
  } else if (warp == kWarpMma) {
    // The WHOLE warp runs this loop with warp-uniform control flow (so descriptors and TMEM addresses
    // stay in uniform registers); one elected lane issues the tcgen05 instructions.  A lane-0-only
    // branch made every operand a divergent value (R2UR / ELECT chains on a single dependent thread).
    const bool leader = elect_one();
    const uint64_t q_desc = make_sw128_desc(smem_u32(sm_q), 16, 1024);
    const uint64_t k_desc0 = make_sw128_desc(smem_u32(sm_k), 16, 1024);
    const uint64_t v_desc0 = make_sw128_desc(smem_u32(sm_v), kKvAtomBytes, 1024);
    constexpr int kMaxKS = (kDPV + 15) / 16;
    auto issue_s = [&](int t) {                  // S(t) = Q K_t^T
      const int st = t & 1;
      mbar_wait(&k_full[st], resident ? 0 : (t >> 1) & 1);
      tc_fence_after();
      const uint64_t kd = k_desc0 + static_cast<uint64_t>((st * kv_bytes) >> 4);
      const uint32_t idesc = resident && t + 1 == n_sub ? p.idesc_s_tail : p.idesc_s;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < kMaxKS; ++ks) {
          if (ks < k_steps_s) {
            const uint32_t qo = ((ks >> 2) * kQAtomBytes + (ks & 3) * 32) >> 4;
            const uint32_t ko = ((ks >> 2) * kKvAtomBytes + (ks & 3) * 32) >> 4;
            umma_bf16_ss(tmem_s, q_desc + qo, kd + ko, idesc, ks != 0);
          }
        }
        umma_commit(s_full);
        if (!resident) umma_commit(&k_empty[st]);
        if (t + 1 == n_sub) umma_commit(q_empty);            // the tile's last S product: Q may be reloaded
      }
      __syncwarp();
    };
    uint32_t n_item = 0;                           // (tile, sub-tile) items so far: the phase of s_full / p_full
    for (int i = 0; i < n_tiles; ++i) {
      mbar_wait(q_full, i & 1);
      issue_s(0);
      for (int t = 0; t < n_sub; ++t, ++n_item) {
        const int st = t & 1;
        ATT_TRACE(kWarpMma, t, 0);
        mbar_wait<64>(p_full, n_item & 1);
        ATT_TRACE(kWarpMma, t, 1);
        mbar_wait(&v_full[st], resident ? 0 : (t >> 1) & 1);
        tc_fence_after();
        // A = P from TMEM: 16 keys = 8 columns of bf16 pairs.  B = V: MN-major (rows = keys), 64-column
        // atoms at LBO = kKvAtomBytes; one K=16 step = 16 key rows = 2048 B.
        const uint64_t vd = v_desc0 + static_cast<uint64_t>((st * kv_bytes) >> 4);
        const int pv_steps = (resident && t + 1 == n_sub ? w_last : kSub) / 16;
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < kSub / 16; ++ks)
            if (ks < pv_steps) umma_bf16_ts(tmem_o, tmem_s + ks * 8, vd + ks * (2048 >> 4), p.idesc_pv, (t | ks) != 0);
          if (!resident) umma_commit(&v_empty[st]);
          if (t + 1 == n_sub) umma_commit(o_done);
        }
        __syncwarp();
        ATT_TRACE(kWarpMma, t, 2);
        if (t + 1 < n_sub) issue_s(t + 1);
        ATT_TRACE(kWarpMma, t, 3);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t t_s = tmem_s + lane_addr;
    const uint32_t t_o = tmem_o + lane_addr;
    float m_run = -INFINITY, l_run = 0.f;
    const float lazy_raw = kLazyLog2 / p.scale_log2;

    auto softmax_sub = [&](int t, auto mask_tag, int valid, int w) {      // w: score columns of this sub-tile (64 / 32)
      constexpr bool kMask = decltype(mask_tag)::value;
      if constexpr (!kMask && kDPV > 64) {
        // Speculative pass, as in attention2_kernel's softmax_sub (two CTAs per SM here: registers for all 64 columns):
        // no row-maximum pass on full sub-tiles after the first; a row sum <= 2^8 proves the reference was still valid
        // and the result is bit-identical to the checked path below.  P is only stored once the check has passed --
        // it overwrites the scores the fallback needs.
        if (t > 0 && w == kSub && p.speculate) {
          const float ms = m_run * p.scale_log2;
          uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
          uint32_t v[kSub], pk[kSub / 2];
          tmem_ld64(t_s, v);
          tmem_ld_wait();
          exp_row<kSub, false>(v, pk, 0, kSub, pack2(p.scale_log2, p.scale_log2), pack2(-ms, -ms), acc);
          const float tot = sum_acc(acc);
          if (!__any_sync(0xffffffffu, !(tot <= 256.0f))) {
            tmem_st32(t_s, pk);
            l_run += tot;
            return;
          }
        }
      }
      // pass 1: row max (two 32-column TMEM loads through the same registers: the CTA must stay under
      // 80 registers per thread for four CTAs per SM)
      float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < kSub; c += 32) {
        if (c >= w) break;
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!kMask || c + i < valid) tm[i & 3] = fmaxf(tm[i & 3], __uint_as_float(v[i]));
      }
      const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
      if (__any_sync(0xffffffffu, tmax > m_run + lazy_raw)) {
        // Rare after the first sub-tile: raise the reference.  S(t) complete implies PV(t-1) complete and
        // PV(t) is not issued before p_full(t), so O is quiescent: rescale it in place.
        const float m_new = fmaxf(m_run, tmax);
        const float alpha = fast_exp2((m_run - m_new) * p.scale_log2);     // 0 on the first sub-tile
        if (t > 0) {
#pragma unroll
          for (int c = 0; c < kDPV; c += 16) {
            uint32_t o[16];
            tmem_ld16(t_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(t_o + c, o);
          }
        }
        l_run *= alpha;
        m_run = m_new;
      }
      // pass 2: P = exp2(S*c - m*c) as bf16 pairs, written over S columns [0, 32): chunk c of S (columns
      // c..c+31) becomes P columns c/2..c/2+15, which only covers S columns this thread has already read.
      const float m_scaled = m_run * p.scale_log2;
      const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2), nms2 = pack2(-m_scaled, -m_scaled);
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int c = 0; c < kSub; c += 32) {
        if (c >= w) break;
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
        exp_row<32, kMask>(v, pk, c, valid, scale2, nms2, acc);
        tmem_st16(t_s + (c >> 1), pk);
      }
      l_run += sum_acc(acc);
    };

    uint32_t n_item = 0;
    for (int i = 0; i < n_tiles; ++i) {
      const int q0 = (tile0 + i) * kBlockQ;
      m_run = -INFINITY;
      l_run = 0.f;
      for (int t = 0; t < n_sub; ++t, ++n_item) {
        int valid = p.seq_k - t * kSub;
        if (p.causal) valid = min(valid, q0 + row - t * kSub + 1);   // per row: keys 0 .. query index
        ATT_TRACE(warp, t, 0);
        mbar_wait<64>(s_full, n_item & 1);
        ATT_TRACE(warp, t, 1);
        tc_fence_after();
        const int w = resident && t + 1 == n_sub ? w_last : kSub;
        if (__all_sync(0xffffffffu, valid >= kSub)) softmax_sub(t, std::false_type{}, kSub, w);
        else softmax_sub(t, std::true_type{}, valid, w);
        ATT_TRACE(warp, t, 2);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        ATT_TRACE(warp, t, 3);
      }
      // O of this tile: its last PV is complete; the next tile's first PV waits for p_full, which these warps only
      // arrive on after the reads below, so the accumulator is not overwritten early.
      mbar_wait<64>(o_done, i & 1);
      tc_fence_after();

      const int s_idx = q0 + row;
      const float inv = 1.0f / l_run;
      __nv_bfloat16* orow = p.o + (static_cast<size_t>(batch) * p.seq_q + s_idx) * p.ld_o + head * p.head_dim;
#pragma unroll
      for (int c = 0; c < kDPV; c += 16) {
        uint32_t v[16];
        tmem_ld16(t_o + c, v);
        tmem_ld_wait();
        if (s_idx < p.seq_q) {
#pragma unroll
          for (int h = 0; h < 16; h += 8) {
            if (c + h < p.head_dim) {
              uint4 u = make_uint4(pack_bf16(__uint_as_float(v[h]) * inv, __uint_as_float(v[h + 1]) * inv),
                                   pack_bf16(__uint_as_float(v[h + 2]) * inv, __uint_as_float(v[h + 3]) * inv),
                                   pack_bf16(__uint_as_float(v[h + 4]) * inv, __uint_as_float(v[h + 5]) * inv),
                                   pack_bf16(__uint_as_float(v[h + 6]) * inv, __uint_as_float(v[h + 7]) * inv));
              *reinterpret_cast<uint4*>(orow + c + h) = u;
            }
          }
        }
      }
      tc_fence_before();                           // the reads above precede this thread's next p_full arrive
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_s);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Two-tile ("ping-pong") variant for head dims <= 64.  ncu on the one-tile kernel at d = 40: 40 % of all issued
// instructions were mbarrier polling (a softmax warp waited for S(t+1) through 17 try_wait rounds per
// sub-tile, at the MUFU pipe's expense: 0.79 IPC per sub-partition, XU 61 % busy).  Here one CTA owns TWO
// 128-query tiles A and B and every softmax thread owns one row of each: while it exponentiates S_B(t) the
// tensor pipe runs PV_A(t) and S_A(t+1), so the next scores are normally complete when the thread comes back and
// the wait is a single successful try_wait.  Two CTAs per SM (2 x 224 TMEM columns, 2 x 97 KB of shared memory):
// 168 registers per thread, so a whole 64-score row is loaded ONCE (tcgen05.ld .x64), reduced, exponentiated and
// written back as 32 packed columns (tcgen05.st .x32); K / V tiles are fetched once per 256 queries through
// 4-stage rings.
//   TMEM columns of tile q:  q*128 + [0, 64) S / P,  q*128 + 64 + [0, kDPV) O.
constexpr int kStages2 = 4;
#ifndef SONIC_ATT_HINT_P
#define SONIC_ATT_HINT_P 64
#endif
#ifndef SONIC_ATT_HINT_S
#define SONIC_ATT_HINT_S 64
#endif
constexpr uint32_t kHintP = SONIC_ATT_HINT_P;   // suspend hints (ns) of the two waits on the S -> softmax -> PV chain of the two-tile kernel
constexpr uint32_t kHintS = SONIC_ATT_HINT_S;

// kResident (cross-attention: at most two key sub-tiles, i.e. <= 128 keys): ONE CTA per (batch, head) keeps K and V
// in shared memory and walks over ALL query-tile pairs of that head, Q double-buffered.  The streaming form launched
// 4096 CTAs of two sub-tiles each for the 64x64 cross-attention layers and spent its time in per-CTA latency
// (TMEM allocation, barrier set-up, the first Q / K / V loads from HBM, the output stores): 126 us for 0.1 ms of
// arithmetic-free work.  Here those costs are paid once per head and the next pair's Q is in flight while the
// current one is processed.
// kNarrow: the last key sub-tile is ragged and spans only 16 or 32 columns (77 keys: 64 + 16).  A template flag because
// the extra selects cost the full-width instantiation 4.7 % (1.433 -> 1.500 ms on the 4096-token self-attention).
// Head dims above 64 can run the same kernel with two / three 64-column shared-memory atoms per Q / K / V tile and 256
// TMEM columns per query tile: ONE CTA per SM (512 columns) whose four softmax warps ping-pong between two tiles.
// That wins where the one-tile kernel only fits one CTA per SM (d = 136 ... 160: the UNet's 16x16 level, 39.6 -> 28.8 us)
// and loses 5 % at d = 80, which keeps the one-tile kernel (attention_plan).
template <int kDPV, bool kResident, bool kNarrow>
__global__ void __launch_bounds__(kAttThreads, (kDPV <= 64 ? 2 : 1))
attention2_kernel(const __grid_constant__ AttParams p) {
  constexpr int kTileCols = kSub + kDPV <= 128 ? 128 : 256;     // TMEM columns of one query tile: S / P, then O
  static_assert(kSub + kDPV <= 256, "S/P + O of one query tile must fit 256 TMEM columns");
  static_assert(kDPV <= 64 || (!kResident && !kNarrow), "head dims > 64: streaming form with full sub-tiles only");
  constexpr int kAtoms = (kDPV + 63) / 64;        // 64-column smem atoms per Q / K / V row (== p.atoms)
  constexpr int kQTile = kAtoms * kQAtomBytes, kKvTile = kAtoms * kKvAtomBytes;
  constexpr int kQStages = kResident ? 2 : 1;     // pairs of Q tiles in flight
  constexpr int kKvStages = kResident ? 2 : (kAtoms == 3 ? 2 : kStages2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sm_q = smem;                                   // kQStages x 2 tiles x kAtoms x 16 KB
  uint8_t* sm_k = sm_q + kQStages * 2 * kQTile;           // kKvStages x kAtoms x 8 KB
  uint8_t* sm_v = sm_k + kKvStages * kKvTile;             // kKvStages x kAtoms x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_v + kKvStages * kKvTile);
  uint64_t* q_full = bars;                        // [2]
  uint64_t* q_empty = bars + 2;                   // [2] every S product of the pair in this Q stage is complete
  uint64_t* k_full = bars + 4;                    // [kStages2]
  uint64_t* k_empty = k_full + kStages2;          // [kStages2]
  uint64_t* v_full = k_empty + kStages2;          // [kStages2]
  uint64_t* v_empty = v_full + kStages2;          // [kStages2]
  uint64_t* s_full = v_empty + kStages2;          // [2] S_q(t) complete (and every earlier MMA, incl. PV_q(t-1))
  uint64_t* p_full = s_full + 2;                  // [2] P_q(t) written over S_q(t), O_q rescaled if it had to be
  uint64_t* o_done = p_full + 2;                  // every PV of the current pair complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_sub = (p.seq_k + kSub - 1) / kSub;
  const int k_steps_s = (p.head_dim + 15) / 16;
  const int total_pairs = (p.seq_q + 2 * kBlockQ - 1) / (2 * kBlockQ);
  const int pair0 = kResident ? 0 : blockIdx.x;                  // this CTA's pairs: pair0, pair0 + 1, ... (n_pairs)
  const int n_pairs = kResident ? total_pairs : 1;
  auto tiles_of = [&](int pair) { return pair * 2 * kBlockQ + kBlockQ < p.seq_q ? 2 : 1; };   // an odd tile count ends in a half pair

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc<2 * kTileCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // the set-up above may overlap the previous kernel's tail (SONIC_PDL)

  if (warp == kWarpTma) {
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v);
      auto load_q = [&](int i) {                   // pair pair0 + i into Q stage i % kQStages
        const int qs = i % kQStages;
        const int nq = tiles_of(pair0 + i);
        if (kResident) mbar_wait<256>(&q_empty[qs], ((i / kQStages) & 1) ^ 1);
        mbar_expect_tx(&q_full[qs], nq * kQTile);
        for (int q = 0; q < nq; ++q)
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            tma_load_4d(sm_q + (qs * 2 + q) * kQTile + a * kQAtomBytes, &p.tm_q, &q_full[qs], a * 64, head,
                        (pair0 + i) * 2 * kBlockQ + q * kBlockQ, batch);
      };
      load_q(0);
      for (int t = 0; t < n_sub; ++t) {
        const int st = t % kKvStages;
        if (!kResident) mbar_wait<512>(&k_empty[st], ((t / kKvStages) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], kKvTile);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_load_4d(sm_k + st * kKvTile + a * kKvAtomBytes, &p.tm_k, &k_full[st], a * 64, head, t * kSub, batch);
        if (!kResident) mbar_wait<512>(&v_empty[st], ((t / kKvStages) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], kKvTile);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          tma_load_4d(sm_v + st * kKvTile + a * kKvAtomBytes, &p.tm_v, &v_full[st], a * 64, head, t * kSub, batch);
      }
      for (int i = 1; i < n_pairs; ++i) load_q(i);
      pdl_launch_dependents();                   // every load of this CTA has been issued
    }
  } else if (warp == kWarpMma) {
    // Warp-uniform control flow, one elected lane issues (see attention_kernel).
    const bool leader = elect_one();
    const uint64_t q_desc0 = make_sw128_desc(smem_u32(sm_q), 16, 1024);
    const uint64_t k_desc0 = make_sw128_desc(smem_u32(sm_k), 16, 1024);
    const uint64_t v_desc0 = make_sw128_desc(smem_u32(sm_v), kKvAtomBytes, 1024);
    constexpr int kMaxKS = (kDPV + 15) / 16;
    auto issue_s = [&](int qs, int q, int st, bool last) {   // S_q = Q_q K^T of the K tile in stage st
      const uint64_t qd = q_desc0 + static_cast<uint64_t>(((qs * 2 + q) * kQTile) >> 4);
      const uint64_t kd = k_desc0 + static_cast<uint64_t>((st * kKvTile) >> 4);
      const uint32_t idesc = kNarrow && last ? p.idesc_s_tail : p.idesc_s;   // a ragged last sub-tile only spans tail_w keys
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < kMaxKS; ++ks)
          if (ks < k_steps_s)                      // K step ks: atom ks / 4, 32-byte slice ks % 4 of its 128-byte rows
            umma_bf16_ss(tmem_base + q * kTileCols, qd + (((ks >> 2) * kQAtomBytes + (ks & 3) * 32) >> 4),
                         kd + (((ks >> 2) * kKvAtomBytes + (ks & 3) * 32) >> 4), idesc, ks != 0);
        umma_commit(&s_full[q]);
      }
      __syncwarp();
    };
    uint32_t n_item = 0;                           // (pair, sub-tile) items so far: the phase of s_full / p_full
    for (int i = 0; i < n_pairs; ++i) {
      const int qs = i % kQStages;
      const int nq = tiles_of(pair0 + i);
      mbar_wait(&q_full[qs], (i / kQStages) & 1);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int q = 0; q < nq; ++q) issue_s(qs, q, 0, n_sub == 1);
      if (leader) {
        if (!kResident) umma_commit(&k_empty[0]);
        if (n_sub == 1) umma_commit(&q_empty[qs]);
      }
      __syncwarp();
      for (int t = 0; t < n_sub; ++t, ++n_item) {
        const int st = t % kKvStages;
        const int st1 = (t + 1) % kKvStages;
        const bool more = t + 1 < n_sub;
        for (int q = 0; q < nq; ++q) {
          ATT_TRACE(kWarpMma, 2 * t + q, 0);
          mbar_wait<kHintP>(&p_full[q], n_item & 1);
          ATT_TRACE(kWarpMma, 2 * t + q, 1);
          if (q == 0) mbar_wait(&v_full[st], kResident ? 0 : (t / kKvStages) & 1);
          tc_fence_after();
          const uint32_t ts = tmem_base + q * kTileCols;
          const uint64_t vd = v_desc0 + static_cast<uint64_t>((st * kKvTile) >> 4);
          const int pv_steps = !kNarrow || more ? kSub / 16 : p.tail_w / 16;
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < kSub / 16; ++ks)
              if (!kNarrow || ks < pv_steps) umma_bf16_ts(ts + kSub, ts + ks * 8, vd + ks * (2048 >> 4), p.idesc_pv, (t | ks) != 0);
            if (!kResident && q == nq - 1) umma_commit(&v_empty[st]);
            if (!more && q == nq - 1) umma_commit(o_done);
          }
          __syncwarp();
          ATT_TRACE(kWarpMma, 2 * t + q, 2);
          if (more) {
            if (q == 0) { mbar_wait(&k_full[st1], kResident ? 0 : ((t + 1) / kKvStages) & 1); tc_fence_after(); }
            issue_s(qs, q, st1, t + 2 == n_sub);
            ATT_TRACE(kWarpMma, 2 * t + q, 3);
            if (q == nq - 1) {
              if (leader) {
                if (!kResident) umma_commit(&k_empty[st1]);
                if (t + 2 == n_sub) umma_commit(&q_empty[qs]);      // the pair's last S product
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float lazy_raw = kLazyLog2 / p.scale_log2;

    auto softmax_sub = [&](float& m_r, float& l_r, uint32_t t_s, int t, auto mask_tag, auto width_tag, int valid) {
      constexpr bool kMask = decltype(mask_tag)::value;
      constexpr int kW = decltype(width_tag)::value;     // score columns of this sub-tile: 64, or a ragged tail's 16 / 32
      uint32_t v[kW];
      if constexpr (kW == 64) tmem_ld64(t_s, v);
      else if constexpr (kW == 32) tmem_ld32(t_s, v);
      else tmem_ld16(t_s, v);
      tmem_ld_wait();
      if constexpr (!kMask && kW == 64) {
        // Speculative pass (full sub-tiles after the first): exponentiate against the CURRENT reference without
        // computing the row maximum.  Any P > 2^8 (the lazy bound) forces the row sum above 2^8, so a sum <= 2^8 proves
        // the reference was still valid and the result is bit-identical to the checked path below; otherwise (rare:
        // the scores of this sub-tile exceed the running maximum) fall through and redo the row the careful way.
        if (t > 0 && p.speculate) {
          const float ms = m_r * p.scale_log2;
          uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
          uint32_t pk[kW / 2];
          exp_row<kW, false>(v, pk, 0, kW, pack2(p.scale_log2, p.scale_log2), pack2(-ms, -ms), acc);
          const float tot = sum_acc(acc);
          if (!__any_sync(0xffffffffu, !(tot <= 256.0f))) {
            tmem_st32(t_s, pk);
            l_r += tot;
            return;
          }
        }
      }
      float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < kW; ++i)
        if (!kMask || i < valid) tm[i & 3] = fmaxf(tm[i & 3], __uint_as_float(v[i]));
      const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
      if (__any_sync(0xffffffffu, tmax > m_r + lazy_raw)) {
        // Rare after the first sub-tile.  S_q(t) complete implies PV_q(t-1) complete and PV_q(t) is not issued
        // before p_full[q](t): O_q is quiescent, rescale it in place.
        const float m_new = fmaxf(m_r, tmax);
        const float alpha = fast_exp2((m_r - m_new) * p.scale_log2);     // 0 on the first sub-tile
        if (t > 0) {
#pragma unroll
          for (int c = 0; c < kDPV; c += 16) {
            uint32_t o[16];
            tmem_ld16(t_s + kSub + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(t_s + kSub + c, o);
          }
        }
        l_r *= alpha;
        m_r = m_new;
      }
      const float m_scaled = m_r * p.scale_log2;
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[kW / 2];
      exp_row<kW, kMask>(v, pk, 0, valid, pack2(p.scale_log2, p.scale_log2), pack2(-m_scaled, -m_scaled), acc);
      if constexpr (kW == 64) tmem_st32(t_s, pk);
      else if constexpr (kW == 32) tmem_st16(t_s, pk);
      else tmem_st8(t_s, pk);
      l_r += sum_acc(acc);
    };

    uint32_t n_item = 0;
    for (int i = 0; i < n_pairs; ++i) {
      const int q0 = (pair0 + i) * 2 * kBlockQ;
      const int nq = tiles_of(pair0 + i);
      float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
      for (int t = 0; t < n_sub; ++t, ++n_item) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q < nq) {
            int valid = p.seq_k - t * kSub;
            if (p.causal) valid = min(valid, q0 + q * kBlockQ + row - t * kSub + 1);
            const uint32_t t_s = tmem_base + q * kTileCols + lane_addr;
            ATT_TRACE(warp, 2 * t + q, 0);
            mbar_wait<kHintS>(&s_full[q], n_item & 1);
            ATT_TRACE(warp, 2 * t + q, 1);
            tc_fence_after();
            using W64 = std::integral_constant<int, 64>;
            const int w = kNarrow && t + 1 == n_sub ? p.tail_w : kSub;
            if (kNarrow && w == 16) softmax_sub(m_run[q], l_run[q], t_s, t, std::true_type{}, std::integral_constant<int, 16>{}, valid);
            else if (kNarrow && w == 32) softmax_sub(m_run[q], l_run[q], t_s, t, std::true_type{}, std::integral_constant<int, 32>{}, valid);
            else if (__all_sync(0xffffffffu, valid >= kSub)) softmax_sub(m_run[q], l_run[q], t_s, t, std::false_type{}, W64{}, kSub);
            else softmax_sub(m_run[q], l_run[q], t_s, t, std::true_type{}, W64{}, valid);
            ATT_TRACE(warp, 2 * t + q, 2);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[q]);
            ATT_TRACE(warp, 2 * t + q, 3);
          }
        }
      }
      // O of this pair: its last PV is complete; the next pair's first PV waits for p_full, which these warps only
      // arrive on after the reads below, so the accumulators are not overwritten early.
      mbar_wait<64>(o_done, i & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q < nq) {
          const int s_idx = q0 + q * kBlockQ + row;
          const float inv = 1.0f / l_run[q];
          __nv_bfloat16* orow = p.o + (static_cast<size_t>(batch) * p.seq_q + s_idx) * p.ld_o + head * p.head_dim;
          const uint32_t t_o = tmem_base + q * kTileCols + kSub + lane_addr;
#pragma unroll
          for (int c = 0; c < kDPV; c += 16) {
            uint32_t v[16];
            tmem_ld16(t_o + c, v);
            tmem_ld_wait();
            if (s_idx < p.seq_q) {
#pragma unroll
              for (int h = 0; h < 16; h += 8) {
                if (c + h < p.head_dim) {
                  uint4 u = make_uint4(pack_bf16(__uint_as_float(v[h]) * inv, __uint_as_float(v[h + 1]) * inv),
                                       pack_bf16(__uint_as_float(v[h + 2]) * inv, __uint_as_float(v[h + 3]) * inv),
                                       pack_bf16(__uint_as_float(v[h + 4]) * inv, __uint_as_float(v[h + 5]) * inv),
                                       pack_bf16(__uint_as_float(v[h + 6]) * inv, __uint_as_float(v[h + 7]) * inv));
                  *reinterpret_cast<uint4*>(orow + c + h) = u;
                }
              }
            }
          }
        }
      }
      tc_fence_before();                           // the reads above precede this thread's next p_full arrive
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<2 * kTileCols>(tmem_base);
  }
}

template <int kDPV, bool kResident, bool kNarrow>
int launch_att2(const AttentionPlan* pl, const AttParams& prm, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SONIC_CUDA(cudaFuncSetAttribute(attention2_kernel<kDPV, kResident, kNarrow>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (pdl_enabled())
    SONIC_CUDA(launch_kernel_ex(attention2_kernel<kDPV, kResident, kNarrow>, pl->grid, dim3(kAttThreads), pl->smem, stream,
                                1, prm));
  else
    attention2_kernel<kDPV, kResident, kNarrow><<<pl->grid, kAttThreads, pl->smem, stream>>>(prm);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

template <int kDPV>
int dispatch_att2(const AttentionPlan* pl, const AttParams& prm, cudaStream_t stream) {
  const bool narrow = prm.tail_w < kSub;
  if constexpr (kDPV > 64) {
    return launch_att2<kDPV, false, false>(pl, prm, stream);      // attention_plan only sends full-sub-tile streaming shapes here
  } else {
    if (pl->resident) return narrow ? launch_att2<kDPV, true, true>(pl, prm, stream) : launch_att2<kDPV, true, false>(pl, prm, stream);
    return narrow ? launch_att2<kDPV, false, true>(pl, prm, stream) : launch_att2<kDPV, false, false>(pl, prm, stream);
  }
}

template <int kDPV>
int launch_att(const AttentionPlan* pl, const AttParams& prm, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<kDPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (pdl_enabled())
    SONIC_CUDA(launch_kernel_ex(attention_kernel<kDPV>, pl->grid, dim3(kAttThreads), pl->smem, stream, 1, prm));
  else
    attention_kernel<kDPV><<<pl->grid, kAttThreads, pl->smem, stream>>>(prm);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int make_qkv_map(CUtensorMap* m, const void* base, int ld, int seq, int batch, int heads, int d, int rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(heads), static_cast<uint64_t>(seq),
                      static_cast<uint64_t>(batch)};
  uint64_t str[3] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(ld) * 2,
                     static_cast<uint64_t>(seq) * ld * 2};
  uint32_t box[4] = {64, 1, static_cast<uint32_t>(rows), 1};
  return encode_tensor_map(m, base, 4, dims, str, box, 128);
}

int round_dpv(int d16) { return d16 <= 48 ? 48 : d16 <= 64 ? 64 : d16 <= 80 ? 80 : d16 <= 128 ? 128 : 160; }

}  // namespace

#ifdef SONIC_ATT_TRACE
}  // namespace sonic
extern "C" int sonic_debug_att_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, sonic::g_att_trace, sizeof(long long) * 8 * 160 * 4) == cudaSuccess ? 0 : -1;
}
namespace sonic {
#endif

double attention_flops(const AttentionOp& op) {
  return 4.0 * op.batch * op.heads * static_cast<double>(op.seq_q) * op.seq_k * op.head_dim;
}

int attention_plan(const AttentionOp& op, AttentionPlan** out) {
  SONIC_REQUIRE(op.q && op.k && op.v && op.o, "attention: null operand");
  SONIC_REQUIRE(op.head_dim % 8 == 0 && op.head_dim >= 16 && op.head_dim <= 160,
                "attention: head_dim=%d unsupported (8 | d, 16 <= d <= 160)", op.head_dim);
  SONIC_REQUIRE(op.seq_q > 0 && op.seq_k > 0, "attention: empty sequence");
  SONIC_REQUIRE(op.ld_q % 8 == 0 && op.ld_k % 8 == 0 && op.ld_v % 8 == 0 && op.ld_o % 8 == 0,
                "attention: row pitches must be multiples of 8 elements");
  auto* pl = new AttentionPlan();
  pl->op = op;
  pl->dpv = round_dpv((op.head_dim + 15) / 16 * 16);
  pl->atoms = (op.head_dim + 63) / 64;
  int rc = make_qkv_map(&pl->tm_q, op.q, op.ld_q, op.seq_q, op.batch, op.heads, op.head_dim, kBlockQ);
  if (!rc) rc = make_qkv_map(&pl->tm_k, op.k, op.ld_k, op.seq_k, op.batch, op.heads, op.head_dim, kSub);
  if (!rc) rc = make_qkv_map(&pl->tm_v, op.v, op.ld_v, op.seq_k, op.batch, op.heads, op.head_dim, kSub);
  if (rc) { delete pl; return rc; }
  pl->qt = (pl->dpv <= 64 && op.seq_q > kBlockQ) ? 2 : 1;
  {
    // Head dims above 64 with whole 64-key sub-tiles: the two-tile kernel with one CTA per SM.  Measured at 32 x 8 heads,
    // 1024 tokens (tools/att_dsweep.py): d = 136 / 160 (three smem atoms: the one-tile kernel fits ONE CTA per SM)
    // 323-326 -> 202 us, d = 96 / 128 188 -> 182 us, but d = 72 / 80 162-166 -> 171-175 us (two one-tile CTAs per SM keep
    // two softmax warps per sub-partition busy, the two-tile CTA only one) -- so d <= 80 stays on the one-tile kernel.
    // SONIC_ATT2_WIDE=0 / 2: never / for every head dim above 64.
    const char* e = getenv("SONIC_ATT2_WIDE");
    const int wide_min = (e && e[0] == '0') ? 1 << 30 : (e && e[0] == '2') ? 65 : 128;
    if (pl->dpv >= wide_min && op.seq_q > kBlockQ && op.seq_k % kSub == 0 && op.seq_k > 2 * kSub && !op.causal) pl->qt = 2;
  }
  // Short key sequences (the cross-attention layers: 77 keys): the one-tile kernel in loop mode -- K / V resident, a
  // CTA walks over several query tiles, four (d <= 64) or two CTAs per SM -- sized so that ONE wave covers the grid.
  const int n_qt = (op.seq_q + kBlockQ - 1) / kBlockQ;
  const char* loop_env = getenv("SONIC_ATT_LOOP");
  if (op.seq_k <= 2 * kSub && n_qt > 1 && !(loop_env && loop_env[0] == '0')) {
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int slots = (pl->dpv <= 64 ? 4 : 2) * sms;
    const int per_bh = std::max(1, std::min(n_qt, slots / std::max(1, op.heads * op.batch)));
    pl->tiles_per_cta = (n_qt + per_bh - 1) / per_bh;
    pl->qt = 1;
  }
  if (pl->qt == 2)
    pl->smem = static_cast<size_t>(pl->atoms) * (2 * kQAtomBytes + 2 * (pl->atoms == 3 ? 2 : kStages2) * kKvAtomBytes) + 1024 + 256;
  else
    pl->smem = static_cast<size_t>(pl->atoms) * (kQAtomBytes + 4 * kKvAtomBytes) + 1024 + 256;
  // <= 128 keys and at least four query-tile pairs per head: one CTA per (batch, head) keeps K / V resident
  pl->resident = pl->qt == 2 && op.seq_k <= 2 * kSub && op.seq_q >= 8 * kBlockQ;
  pl->grid = dim3(pl->resident ? 1 : (op.seq_q + kBlockQ * pl->qt - 1) / (kBlockQ * pl->qt), op.heads, op.batch);
  if (pl->qt == 1) pl->grid.x = (n_qt + pl->tiles_per_cta - 1) / pl->tiles_per_cta;
  *out = pl;
  return 0;
}

void attention_plan_free(AttentionPlan* plan) { delete plan; }

double attention_plan_flops(const AttentionPlan* plan) { return plan ? attention_flops(plan->op) : 0.0; }

int attention_launch(const AttentionPlan* pl, cudaStream_t stream) {
  const AttentionOp& op = pl->op;
  AttParams prm;
  prm.tm_q = pl->tm_q; prm.tm_k = pl->tm_k; prm.tm_v = pl->tm_v;
  prm.o = static_cast<__nv_bfloat16*>(op.o);
  prm.ld_o = op.ld_o; prm.seq_q = op.seq_q; prm.seq_k = op.seq_k; prm.head_dim = op.head_dim;
  prm.atoms = pl->atoms;
  prm.causal = op.causal;
  prm.scale_log2 = op.scale * 1.4426950408889634f;
  prm.idesc_s = make_idesc_bf16(kBlockQ, kSub, false);
  const int tail = op.seq_k - (op.seq_k - 1) / kSub * kSub;          // 1 .. 64 keys in the last sub-tile
  prm.tail_w = tail <= 16 ? 16 : tail <= 32 ? 32 : kSub;
  prm.idesc_s_tail = make_idesc_bf16(kBlockQ, prm.tail_w, false);
  prm.idesc_pv = make_idesc_bf16(kBlockQ, pl->dpv, true);
  prm.tiles_per_cta = pl->tiles_per_cta;
  {
    const char* e = getenv("SONIC_ATT_SPEC");                         // "0": always take the checked path (A/B, tests)
    prm.speculate = (e && e[0] == '0') ? 0 : 1;
  }
  if (pl->qt == 2) {
    switch (pl->dpv) {
      case 48: return dispatch_att2<48>(pl, prm, stream);
      case 64: return dispatch_att2<64>(pl, prm, stream);
      case 80: return dispatch_att2<80>(pl, prm, stream);
      case 128: return dispatch_att2<128>(pl, prm, stream);
      default: return dispatch_att2<160>(pl, prm, stream);
    }
  }
  switch (pl->dpv) {
    case 48: return launch_att<48>(pl, prm, stream);
    case 64: return launch_att<64>(pl, prm, stream);
    case 80: return launch_att<80>(pl, prm, stream);
    case 128: return launch_att<128>(pl, prm, stream);
    default: return launch_att<160>(pl, prm, stream);
  }
}

}  // namespace sonic
