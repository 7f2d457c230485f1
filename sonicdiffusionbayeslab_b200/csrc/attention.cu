// Flash attention on tcgen05 tensor cores for the UNet's self- and cross-attention layers.
//   O[b, s, h, :] = softmax(scale * Q[b, s, h, :] K[b, :, h, :]^T) V[b, :, h, :]
// Q/K/V are read in place from the projection GEMM outputs ([B*S, ld] rows, head h at column
// h*D) through strided 4-D TMA tensor maps; head dims that are not a multiple of 64 (40, 80,
// 160) are zero-filled by the TMA unit, never padded in HBM.
//
// One CTA = one 128-row query tile of one (batch, head); TWO CTAs are resident per SM (each owns
// 256 of the 512 TMEM columns and <= 113 KB of shared memory) so that one CTA's softmax overlaps
// the other's MMAs and TMEM round trips.  Warp roles (192 threads):
//   warp 0     TMA producer : Q once, then K (2-stage ring) and V (1 stage) tiles of kBKV keys
//   warp 1     MMA issuer   : S = Q K^T  (M128 x N=kBKV, K = D) into TMEM, then
//                             PV = P V   (M128 x N=D, K = kBKV) into a second TMEM region
//   warps 2-5  softmax      : one query row per thread (TMEM lane == row, so row max / sum need
//                             no shuffles): S -> online softmax -> P (bf16, 128B-swizzled smem,
//                             the A operand of the PV MMA); O accumulates in registers.
#include "ops.cuh"

#include <type_traits>

namespace sonic {

struct AttentionPlan {
  CUtensorMap tm_q, tm_k, tm_v;
  AttentionOp op;
  int dpv = 0;         // head dim rounded up to a supported MMA N of the PV product
  int atoms = 0;       // 64-column smem atoms per row (1, 2, 3)
  int bkv = 0;         // keys per tile (128 for one-atom heads, 64 otherwise)
  size_t smem = 0;
  dim3 grid;
};

namespace {

constexpr int kAttThreads = 192;
constexpr int kBlockQ = 128;
constexpr int kQAtomBytes = kBlockQ * 128;      // 128 rows x 64 bf16
constexpr int kTmemCols = 256;

struct AttParams {
  CUtensorMap tm_q, tm_k, tm_v;
  __nv_bfloat16* o;
  int ld_o, seq_q, seq_k, head_dim, atoms;
  float scale_log2;
  uint32_t idesc_s, idesc_pv;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kDPV, int kBKV>
__global__ void __launch_bounds__(kAttThreads, 2)
attention_kernel(const __grid_constant__ AttParams p) {
  constexpr int kKvAtomBytes = kBKV * 128;       // kBKV rows x 64 bf16
  constexpr int kPAtoms = kBKV / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int q_bytes = p.atoms * kQAtomBytes;
  const int kv_bytes = p.atoms * kKvAtomBytes;
  uint8_t* sm_q = smem;
  uint8_t* sm_k = sm_q + q_bytes;                // 2 stages
  uint8_t* sm_v = sm_k + 2 * kv_bytes;           // 1 stage
  uint8_t* sm_p = sm_v + kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + kPAtoms * kQAtomBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;        // P written AND S consumed
  uint64_t* p_empty = bars + 9;
  uint64_t* o_full = bars + 10;
  uint64_t* o_empty = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBlockQ;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.seq_k + kBKV - 1) / kBKV;
  const int k_steps_s = (p.head_dim + 15) / 16;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;
  const uint32_t tmem_o = tmem_s + kBKV;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v);
      mbar_expect_tx(q_full, q_bytes);
      for (int a = 0; a < p.atoms; ++a)
        tma_load_4d(sm_q + a * kQAtomBytes, &p.tm_q, q_full, a * 64, head, q0, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], kv_bytes);
        for (int a = 0; a < p.atoms; ++a)
          tma_load_4d(sm_k + st * kv_bytes + a * kKvAtomBytes, &p.tm_k, &k_full[st], a * 64, head, j * kBKV, batch);
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_expect_tx(v_full, kv_bytes);
        for (int a = 0; a < p.atoms; ++a)
          tma_load_4d(sm_v + a * kKvAtomBytes, &p.tm_v, v_full, a * 64, head, j * kBKV, batch);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      auto issue_s = [&](int j) {              // S(j) = Q K_j^T ; the S region is free when called
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sm_q), ka = smem_u32(sm_k + st * kv_bytes);
        for (int ks = 0; ks < k_steps_s; ++ks) {
          const uint64_t da = make_sw128_desc(qa + (ks >> 2) * kQAtomBytes + (ks & 3) * 32, 16, 1024);
          const uint64_t db = make_sw128_desc(ka + (ks >> 2) * kKvAtomBytes + (ks & 3) * 32, 16, 1024);
          umma_bf16_ss(tmem_s, da, db, p.idesc_s, ks != 0);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(p_full, j & 1);              // P(j) in smem, S(j) fully read by the softmax warps
        if (j + 1 < n_kv) issue_s(j + 1);      // queue the next scores first: softmax restarts sooner
        mbar_wait(v_full, j & 1);
        mbar_wait(o_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sm_p), va = smem_u32(sm_v);
        for (int ks = 0; ks < kBKV / 16; ++ks) {
          // A = P: K-major, 64-key atoms.  B = V: MN-major (rows = keys), 64-column atoms at
          // LBO = kKvAtomBytes; one K=16 step = 16 key rows = 2048 B.
          const uint64_t da = make_sw128_desc(pa + (ks >> 2) * kQAtomBytes + (ks & 3) * 32, 16, 1024);
          const uint64_t db = make_sw128_desc(va + ks * 2048, kKvAtomBytes, 1024);
          umma_bf16_ss(tmem_o, da, db, p.idesc_pv, ks != 0);
        }
        umma_commit(o_full);
        umma_commit(v_empty);
        umma_commit(p_empty);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_s + lane_addr;
    float o_acc[kDPV];
#pragma unroll
    for (int i = 0; i < kDPV; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto accumulate = [&](int j, float alpha) {
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kDPV; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_o + lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    };

    // one key tile: row max, then exponentials -> P (bf16) in smem; kMask only on a ragged last tile
    auto softmax_tile = [&](int j, auto mask_tag, int valid) -> float {
      constexpr bool kMask = decltype(mask_tag)::value;
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < kBKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!kMask || c + i < valid) tmax = fmaxf(tmax, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, tmax);
      const float alpha = fast_exp2((m_run - m_new) * p.scale_log2);
      const float m_scaled = m_new * p.scale_log2;
      mbar_wait(p_empty, (j & 1) ^ 1);
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < kBKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_scaled));
          if (kMask && c + i >= valid) e[i] = 0.f;
          psum += e[i];
        }
        // 128B-swizzled K-major store: 16-byte chunk index XOR (row & 7)
        uint8_t* prow = sm_p + (c >> 6) * kQAtomBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 63) >> 3) + q;
          uint4 u = make_uint4(pack_bf16(e[8 * q], e[8 * q + 1]), pack_bf16(e[8 * q + 2], e[8 * q + 3]),
                               pack_bf16(e[8 * q + 4], e[8 * q + 5]), pack_bf16(e[8 * q + 6], e[8 * q + 7]));
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = u;
        }
      }
      l_run = fmaf(l_run, alpha, psum);
      m_run = m_new;
      return alpha;
    };

    for (int j = 0; j < n_kv; ++j) {
      const int valid = p.seq_k - j * kBKV;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const float alpha = valid >= kBKV ? softmax_tile(j, std::false_type{}, kBKV)
                                        : softmax_tile(j, std::true_type{}, valid);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (j > 0) accumulate(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate(n_kv - 1, alpha_prev);

    const int s_idx = q0 + row;
    if (s_idx < p.seq_q) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* orow = p.o + (static_cast<size_t>(batch) * p.seq_q + s_idx) * p.ld_o + head * p.head_dim;
#pragma unroll
      for (int c = 0; c < kDPV; c += 8) {
        if (c < p.head_dim) {
          uint4 u = make_uint4(pack_bf16(o_acc[c] * inv, o_acc[c + 1] * inv),
                               pack_bf16(o_acc[c + 2] * inv, o_acc[c + 3] * inv),
                               pack_bf16(o_acc[c + 4] * inv, o_acc[c + 5] * inv),
                               pack_bf16(o_acc[c + 6] * inv, o_acc[c + 7] * inv));
          *reinterpret_cast<uint4*>(orow + c) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_s);
  }
}

template <int kDPV, int kBKV>
int launch_att(const AttentionPlan* pl, const AttParams& prm, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<kDPV, kBKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    attr_set = true;
  }
  attention_kernel<kDPV, kBKV><<<pl->grid, kAttThreads, pl->smem, stream>>>(prm);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int make_qkv_map(CUtensorMap* m, const void* base, int ld, int seq, int batch, int heads, int d, int rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(heads), static_cast<uint64_t>(seq),
                      static_cast<uint64_t>(batch)};
  uint64_t str[3] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(ld) * 2,
                     static_cast<uint64_t>(seq) * ld * 2};
  uint32_t box[4] = {64, 1, static_cast<uint32_t>(rows), 1};
  return encode_tensor_map(m, base, 4, dims, str, box, true);
}

int round_dpv(int d16) { return d16 <= 48 ? 48 : d16 <= 64 ? 64 : d16 <= 80 ? 80 : d16 <= 128 ? 128 : 160; }

}  // namespace

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
  pl->bkv = pl->atoms == 1 ? 128 : 64;
  int rc = make_qkv_map(&pl->tm_q, op.q, op.ld_q, op.seq_q, op.batch, op.heads, op.head_dim, kBlockQ);
  if (!rc) rc = make_qkv_map(&pl->tm_k, op.k, op.ld_k, op.seq_k, op.batch, op.heads, op.head_dim, pl->bkv);
  if (!rc) rc = make_qkv_map(&pl->tm_v, op.v, op.ld_v, op.seq_k, op.batch, op.heads, op.head_dim, pl->bkv);
  if (rc) { delete pl; return rc; }
  const size_t kv = static_cast<size_t>(pl->atoms) * pl->bkv * 128;
  pl->smem = static_cast<size_t>(pl->atoms) * kQAtomBytes + 3 * kv + (pl->bkv / 64) * kQAtomBytes + 1024 + 256;
  pl->grid = dim3((op.seq_q + kBlockQ - 1) / kBlockQ, op.heads, op.batch);
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
  prm.scale_log2 = op.scale * 1.4426950408889634f;
  prm.idesc_s = make_idesc_bf16(kBlockQ, pl->bkv, false);
  prm.idesc_pv = make_idesc_bf16(kBlockQ, pl->dpv, true);
  if (pl->bkv == 128) {
    if (pl->dpv == 48) return launch_att<48, 128>(pl, prm, stream);
    return launch_att<64, 128>(pl, prm, stream);
  }
  switch (pl->dpv) {
    case 80: return launch_att<80, 64>(pl, prm, stream);
    case 128: return launch_att<128, 64>(pl, prm, stream);
    default: return launch_att<160, 64>(pl, prm, stream);
  }
}

}  // namespace sonic
