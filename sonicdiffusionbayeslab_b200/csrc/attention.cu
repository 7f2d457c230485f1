// Flash attention on tcgen05 tensor cores for the UNet's self- and cross-attention layers.
//   O[b, s, h, :] = softmax(scale * Q[b, s, h, :] K[b, :, h, :]^T) V[b, :, h, :]
// Q/K/V are read in place from the projection GEMM outputs ([B*S, ld] rows, head h at column
// h*D) through strided 4-D TMA tensor maps; head dims that are not a multiple of 64 (40, 80,
// 160) are zero-filled by the TMA unit, never padded in HBM.
//
// One CTA = one 128-row query tile of one (batch, head).  Warp roles (192 threads):
//   warp 0     TMA producer : Q once, then K / V tiles (128 keys) through separate rings
//   warp 1     MMA issuer   : S = Q K^T  (M128 x N128, K = D) into a double-buffered TMEM tile,
//                             then  PV = P V (M128 x N=D, K = 128) into a TMEM scratch tile
//   warps 2-5  softmax      : one query row per thread (TMEM lane == row, so row max / sum need
//                             no shuffles): S -> online softmax -> P (bf16, 128B-swizzled smem,
//                             the A operand of the PV MMA); O accumulates in registers.
#include "ops.cuh"

namespace sonic {

struct AttentionPlan {
  CUtensorMap tm_q, tm_k, tm_v;
  AttentionOp op;
  int dpv = 0;         // head dim rounded up to 16 (MMA N of the PV product)
  int atoms = 0;       // 64-column smem atoms per row (1, 2, 3)
  int stages = 0;
  size_t smem = 0;
  dim3 grid;
};

namespace {

constexpr int kAttThreads = 192;
constexpr int kBlockQ = 128;
constexpr int kBlockKV = 128;
constexpr int kAtomBytes = 128 * 128;      // 128 rows x 64 bf16

struct AttParams {
  CUtensorMap tm_q, tm_k, tm_v;
  __nv_bfloat16* o;
  int ld_o, seq_q, seq_k, head_dim, atoms, stages;
  float scale_log2;
  uint32_t idesc_s, idesc_pv;
};

template <int kDPV>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_kernel(const __grid_constant__ AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int tile_bytes = p.atoms * kAtomBytes;
  uint8_t* sm_q = smem;
  uint8_t* sm_k = sm_q + tile_bytes;
  uint8_t* sm_v = sm_k + p.stages * tile_bytes;
  uint8_t* sm_p = sm_v + p.stages * tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + 2 * kAtomBytes);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // [2]
  uint64_t* s_empty = bars + 11;      // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* p_empty = bars + 14;
  uint64_t* o_full = bars + 15;
  uint64_t* o_empty = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBlockQ;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.seq_k + kBlockKV - 1) / kBlockKV;
  const int k_steps_s = (p.head_dim + 15) / 16;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * kBlockKV;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_q); tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v);
      mbar_expect_tx(q_full, tile_bytes);
      for (int a = 0; a < p.atoms; ++a)
        tma_load_4d(sm_q + a * kAtomBytes, &p.tm_q, q_full, a * 64, head, q0, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % p.stages;
        const uint32_t ph = (j / p.stages) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], tile_bytes);
        for (int a = 0; a < p.atoms; ++a)
          tma_load_4d(sm_k + st * tile_bytes + a * kAtomBytes, &p.tm_k, &k_full[st], a * 64, head,
                      j * kBlockKV, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], tile_bytes);
        for (int a = 0; a < p.atoms; ++a)
          tma_load_4d(sm_v + st * tile_bytes + a * kAtomBytes, &p.tm_v, &v_full[st], a * 64, head,
                      j * kBlockKV, batch);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      auto issue_s = [&](int j) {
        const int st = j % p.stages;
        const int b = j & 1;
        mbar_wait(&k_full[st], (j / p.stages) & 1);
        mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sm_q), ka = smem_u32(sm_k + st * tile_bytes);
        for (int ks = 0; ks < k_steps_s; ++ks) {
          const uint32_t off = (ks >> 2) * kAtomBytes + (ks & 3) * 32;
          umma_bf16_ss(tmem_base + b * kBlockKV, make_sw128_desc(qa + off, 16, 1024),
                       make_sw128_desc(ka + off, 16, 1024), p.idesc_s, ks != 0);
        }
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j % p.stages;
        mbar_wait(&v_full[st], (j / p.stages) & 1);
        mbar_wait(p_full, j & 1);
        mbar_wait(o_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sm_p), va = smem_u32(sm_v + st * tile_bytes);
        for (int ks = 0; ks < kBlockKV / 16; ++ks) {
          // A = P: K-major, 64-key atoms.  B = V: MN-major (rows = keys), 64-column atoms at
          // LBO = kAtomBytes; one K=16 step = 16 key rows = 2048 B.
          const uint64_t da = make_sw128_desc(pa + (ks >> 2) * kAtomBytes + (ks & 3) * 32, 16, 1024);
          const uint64_t db = make_sw128_desc(va + ks * 2048, kAtomBytes, 1024);
          umma_bf16_ss(tmem_o, da, db, p.idesc_pv, ks != 0);
        }
        umma_commit(o_full);
        umma_commit(&v_empty[st]);
        umma_commit(p_empty);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
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
        for (int i = 0; i < 16; ++i) o_acc[c + i] = o_acc[c + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const int valid = min(kBlockKV, p.seq_k - j * kBlockKV);
      const uint32_t t_s = tmem_base + lane_addr + b * kBlockKV;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < kBlockKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid) tmax = fmaxf(tmax, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, tmax);
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);
      const float m_scaled = m_new * p.scale_log2;
      mbar_wait(p_empty, (j & 1) ^ 1);
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < kBlockKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + c, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = c + i < valid ? exp2f(__uint_as_float(v[i]) * p.scale_log2 - m_scaled) : 0.f;
          psum += e[i];
        }
        // 128B-swizzled K-major store: 16-byte chunk index XOR (row & 7)
        uint8_t* prow = sm_p + (c >> 6) * kAtomBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 63) >> 3) + q;
          uint4 u = make_uint4(pack_bf16(e[8 * q], e[8 * q + 1]), pack_bf16(e[8 * q + 2], e[8 * q + 3]),
                               pack_bf16(e[8 * q + 4], e[8 * q + 5]), pack_bf16(e[8 * q + 6], e[8 * q + 7]));
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = u;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&s_empty[b]); mbar_arrive(p_full); }
      l_run = l_run * alpha + psum;
      m_run = m_new;
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
    tmem_dealloc<512>(tmem_base);
  }
}

bool g_att_attr_set = false;

template <int kDPV>
int launch_att(const AttentionPlan* pl, const AttParams& prm, cudaStream_t stream) {
  attention_kernel<kDPV><<<pl->grid, kAttThreads, pl->smem, stream>>>(prm);
  SONIC_CUDA(cudaGetLastError());
  return 0;
}

int make_qkv_map(CUtensorMap* m, const void* base, int ld, int seq, int batch, int heads, int d) {
  uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(heads), static_cast<uint64_t>(seq),
                      static_cast<uint64_t>(batch)};
  uint64_t str[3] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(ld) * 2,
                     static_cast<uint64_t>(seq) * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return encode_tensor_map(m, base, 4, dims, str, box, true);
}

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
  pl->dpv = (op.head_dim + 15) / 16 * 16;
  pl->atoms = (op.head_dim + 63) / 64;
  pl->stages = pl->atoms <= 2 ? 2 : 1;
  int rc = make_qkv_map(&pl->tm_q, op.q, op.ld_q, op.seq_q, op.batch, op.heads, op.head_dim);
  if (!rc) rc = make_qkv_map(&pl->tm_k, op.k, op.ld_k, op.seq_k, op.batch, op.heads, op.head_dim);
  if (!rc) rc = make_qkv_map(&pl->tm_v, op.v, op.ld_v, op.seq_k, op.batch, op.heads, op.head_dim);
  if (rc) { delete pl; return rc; }
  const size_t tiles = static_cast<size_t>(pl->atoms) * kAtomBytes;
  pl->smem = tiles * (1 + 2 * pl->stages) + 2 * kAtomBytes + 1024 + 256;
  if (pl->smem < 120 * 1024) pl->smem = 120 * 1024;   // one CTA per SM: it owns all 512 TMEM columns
  pl->grid = dim3((op.seq_q + kBlockQ - 1) / kBlockQ, op.heads, op.batch);
  *out = pl;
  return 0;
}

void attention_plan_free(AttentionPlan* plan) { delete plan; }

double attention_plan_flops(const AttentionPlan* plan) { return plan ? attention_flops(plan->op) : 0.0; }

int attention_launch(const AttentionPlan* pl, cudaStream_t stream) {
  if (!g_att_attr_set) {
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SONIC_CUDA(cudaFuncSetAttribute(attention_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_att_attr_set = true;
  }
  const AttentionOp& op = pl->op;
  AttParams prm;
  prm.tm_q = pl->tm_q; prm.tm_k = pl->tm_k; prm.tm_v = pl->tm_v;
  prm.o = static_cast<__nv_bfloat16*>(op.o);
  prm.ld_o = op.ld_o; prm.seq_q = op.seq_q; prm.seq_k = op.seq_k; prm.head_dim = op.head_dim;
  prm.atoms = pl->atoms; prm.stages = pl->stages;
  prm.scale_log2 = op.scale * 1.4426950408889634f;
  prm.idesc_s = make_idesc_bf16(kBlockQ, kBlockKV, false);
  int dpv = pl->dpv <= 48 ? 48 : pl->dpv <= 64 ? 64 : pl->dpv <= 80 ? 80 : pl->dpv <= 128 ? 128 : 160;
  prm.idesc_pv = make_idesc_bf16(kBlockQ, dpv, true);
  switch (dpv) {
    case 48: return launch_att<48>(pl, prm, stream);
    case 64: return launch_att<64>(pl, prm, stream);
    case 80: return launch_att<80>(pl, prm, stream);
    case 128: return launch_att<128>(pl, prm, stream);
    default: return launch_att<160>(pl, prm, stream);
  }
}

}  // namespace sonic
