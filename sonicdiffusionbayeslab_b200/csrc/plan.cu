// Launch-plan executor: the native runtime under the sampling engine.
//
// The Python host walks the network once and records every operator (with its final device
// pointers) into a plan; TMA descriptors, tile shapes and job tables are resolved at record
// time.  Running a plan is then a tight native loop of kernel launches on the caller's stream
// -- or, after sonic_plan_capture(), the replay of one CUDA graph -- with no Python, no
// allocation and no host synchronisation inside.  One plan = one UNet forward variant (full
// step, DeepCache cached step, ...).
#include "../../include/sonic.h"

#include "gemm.cuh"
#include "ops.cuh"

#include <memory>
#include <vector>

namespace sonic {

struct PlanOp {
  enum Kind { kGemm, kAttention, kGroupNorm, kLayerNorm, kToNhwc8, kToNchw, kUpsample, kIm2col, kTimeEmb, kGemv, kSoftmaxRows, kLnSide };
  Kind kind;
  GemmPlan gemm;
  AttentionPlan* att = nullptr;
  GroupNormOp gn;
  // generic small-op arguments
  const void* src = nullptr; void* dst = nullptr;
  const float* f0 = nullptr; const float* f1 = nullptr;
  int i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0;
  long l0 = 0;
  float eps = 0.f;
  GemvJob* jobs_dev = nullptr;
};

struct Plan {
  std::vector<PlanOp> ops;
  double flops = 0;
  int launches = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  ~Plan() {
    for (auto& op : ops) {
      if (op.att) attention_plan_free(op.att);
      if (op.jobs_dev) cudaFree(op.jobs_dev);
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
};

static int run_ops(const Plan& plan, cudaStream_t s) {
  for (const PlanOp& op : plan.ops) {
    int rc = 0;
    switch (op.kind) {
      case PlanOp::kGemm: rc = gemm_launch(op.gemm, s); break;
      case PlanOp::kAttention: rc = attention_launch(op.att, s); break;
      case PlanOp::kGroupNorm: rc = groupnorm_launch(op.gn, s); break;
      case PlanOp::kLayerNorm: rc = layernorm_launch(op.src, op.dst, op.i0, op.i1, op.eps, op.f0, op.f1, s); break;
      case PlanOp::kToNhwc8: rc = nchw_to_nhwc8_launch(op.src, op.i0, op.i1, op.i2, op.i3, op.i4, op.dst, s); break;
      case PlanOp::kToNchw: rc = nhwc_to_nchw_launch(op.src, op.i0, op.i1, op.i2, op.i3, op.dst, op.i4, s); break;
      case PlanOp::kUpsample: rc = upsample2x_launch(op.src, op.dst, op.i0, op.i1, op.i2, op.i3, s); break;
      case PlanOp::kIm2col: rc = im2col3x3_launch(op.src, op.dst, op.i0, op.i1, op.i2, op.i3, op.i4, s); break;
      case PlanOp::kTimeEmb: rc = timestep_embedding_launch(op.f0, op.i0, static_cast<float*>(op.dst), s); break;
      case PlanOp::kGemv: rc = gemv_batched_launch(op.jobs_dev, op.i0, op.i1, op.f0, op.i2, op.i3, s); break;
      case PlanOp::kSoftmaxRows: rc = softmax_rows_launch(op.dst, op.i0, op.i1, op.l0, op.eps, s); break;
      case PlanOp::kLnSide:
        rc = ln_side_launch(op.f0, op.i0, op.i1, op.i2, op.eps, op.dst, const_cast<float*>(op.f1), s);
        break;
    }
    if (rc) return rc;
  }
  return 0;
}

}  // namespace sonic

using namespace sonic;

extern "C" {

int sonic_plan_create(sonic_plan_t* out) {
  SONIC_REQUIRE(out != nullptr, "sonic_plan_create: null out");
  *out = new Plan();
  return 0;
}

int sonic_plan_destroy(sonic_plan_t h) {
  delete static_cast<Plan*>(h);
  return 0;
}

int sonic_plan_add_conv_gemm(sonic_plan_t h, const sonic_gemm_args* a) {
  SONIC_REQUIRE(h && a, "sonic_plan_add_conv_gemm: null argument");
  Plan* plan = static_cast<Plan*>(h);
  GemmOp op;
  op.a0 = a->a0; op.c0 = a->c0; op.ld0 = a->ld0;
  op.a1 = a->a1; op.c1 = a->c1; op.ld1 = a->ld1;
  op.n_img = a->n_img; op.H = a->H; op.W = a->W;
  op.w = a->w; op.N = a->N; op.taps = a->taps;
  op.bias = a->bias; op.row_bias = a->row_bias;
  op.residual = a->residual; op.ld_res = a->ld_res;
  op.out = a->out; op.ld_out = a->ld_out;
  op.epilogue = a->epilogue; op.block_n = a->block_n; op.gn_partial = a->gn_partial;
  op.ln_stats_out = a->ln_stats_out; op.row_scale = a->row_scale;
  op.stride = a->stride == 2 ? 2 : 1; op.upsample = a->upsample;
  PlanOp p;
  p.kind = PlanOp::kGemm;
  if (int rc = gemm_plan(op, &p.gemm)) return rc;
  plan->flops += p.gemm.flops;
  plan->launches += 1;
  plan->ops.push_back(p);
  return 0;
}

int sonic_plan_add_attention(sonic_plan_t h, const sonic_attention_args* a) {
  SONIC_REQUIRE(h && a, "sonic_plan_add_attention: null argument");
  Plan* plan = static_cast<Plan*>(h);
  AttentionOp op;
  op.q = a->q; op.k = a->k; op.v = a->v; op.o = a->o;
  op.ld_q = a->ld_q; op.ld_k = a->ld_k; op.ld_v = a->ld_v; op.ld_o = a->ld_o;
  op.batch = a->batch; op.heads = a->heads; op.seq_q = a->seq_q; op.seq_k = a->seq_k;
  op.head_dim = a->head_dim; op.scale = a->scale; op.causal = a->causal;
  PlanOp p;
  p.kind = PlanOp::kAttention;
  if (int rc = attention_plan(op, &p.att)) return rc;
  plan->flops += attention_flops(op);
  plan->launches += 1;
  plan->ops.push_back(p);
  return 0;
}

int sonic_plan_add_groupnorm(sonic_plan_t h, const void* x0, int32_t c0, const void* x1, int32_t c1,
                             int32_t n_img, int32_t hw, int32_t groups, float eps, const float* gamma,
                             const float* beta, int32_t silu, float* stats, void* y) {
  SONIC_REQUIRE(h && x0 && y && stats, "sonic_plan_add_groupnorm: null argument");
  Plan* plan = static_cast<Plan*>(h);
  PlanOp p;
  p.kind = PlanOp::kGroupNorm;
  p.gn.x0 = x0; p.gn.c0 = c0; p.gn.x1 = x1; p.gn.c1 = c1;
  p.gn.n_img = n_img; p.gn.hw = hw; p.gn.groups = groups; p.gn.eps = eps;
  p.gn.gamma = gamma; p.gn.beta = beta; p.gn.silu = silu; p.gn.stats = stats; p.gn.y = y;
  plan->launches += groupnorm_launch_count(p.gn);
  plan->ops.push_back(p);
  return 0;
}

int sonic_plan_add_groupnorm_fused(sonic_plan_t h, const void* x0, int32_t c0, const float* part0, const void* x1,
                                   int32_t c1, const float* part1, int32_t n_img, int32_t hw, int32_t groups,
                                   float eps, const float* gamma, const float* beta, int32_t silu, float* stats,
                                   void* y) {
  SONIC_REQUIRE(h && x0 && y && stats && part0 && (x1 == nullptr || part1 != nullptr),
                "sonic_plan_add_groupnorm_fused: null argument");
  Plan* plan = static_cast<Plan*>(h);
  PlanOp p;
  p.kind = PlanOp::kGroupNorm;
  p.gn.x0 = x0; p.gn.c0 = c0; p.gn.x1 = x1; p.gn.c1 = c1; p.gn.part0 = part0; p.gn.part1 = part1;
  p.gn.n_img = n_img; p.gn.hw = hw; p.gn.groups = groups; p.gn.eps = eps;
  p.gn.gamma = gamma; p.gn.beta = beta; p.gn.silu = silu; p.gn.stats = stats; p.gn.y = y;
  plan->launches += groupnorm_launch_count(p.gn);
  plan->ops.push_back(p);
  return 0;
}

int sonic_plan_add_layernorm(sonic_plan_t h, const void* x, void* y, int32_t rows, int32_t C, float eps,
                             const float* gamma, const float* beta) {
  SONIC_REQUIRE(h && x && y, "sonic_plan_add_layernorm: null argument");
  Plan* plan = static_cast<Plan*>(h);
  PlanOp p;
  p.kind = PlanOp::kLayerNorm;
  p.src = x; p.dst = y; p.i0 = rows; p.i1 = C; p.eps = eps; p.f0 = gamma; p.f1 = beta;
  plan->launches += 1;
  plan->ops.push_back(p);
  return 0;
}

int sonic_plan_add_nchw_to_nhwc8(sonic_plan_t h, const void* x, int32_t dtype, int32_t n_img, int32_t C,
                                 int32_t hw, int32_t dup, void* y) {
  SONIC_REQUIRE(h && x && y, "sonic_plan_add_nchw_to_nhwc8: null argument");
  PlanOp p;
  p.kind = PlanOp::kToNhwc8;
  p.src = x; p.dst = y; p.i0 = dtype; p.i1 = n_img; p.i2 = C; p.i3 = hw; p.i4 = dup;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_nhwc_to_nchw(sonic_plan_t h, const void* x, int32_t ld, int32_t n_img, int32_t C, int32_t hw,
                                void* y, int32_t dtype) {
  SONIC_REQUIRE(h && x && y, "sonic_plan_add_nhwc_to_nchw: null argument");
  PlanOp p;
  p.kind = PlanOp::kToNchw;
  p.src = x; p.dst = y; p.i0 = ld; p.i1 = n_img; p.i2 = C; p.i3 = hw; p.i4 = dtype;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_upsample2x(sonic_plan_t h, const void* x, void* y, int32_t n_img, int32_t H, int32_t W,
                              int32_t C) {
  SONIC_REQUIRE(h && x && y, "sonic_plan_add_upsample2x: null argument");
  PlanOp p;
  p.kind = PlanOp::kUpsample;
  p.src = x; p.dst = y; p.i0 = n_img; p.i1 = H; p.i2 = W; p.i3 = C;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_im2col3x3(sonic_plan_t h, const void* x, void* y, int32_t n_img, int32_t H, int32_t W,
                             int32_t C, int32_t stride) {
  SONIC_REQUIRE(h && x && y && (stride == 1 || stride == 2), "sonic_plan_add_im2col3x3: bad argument");
  PlanOp p;
  p.kind = PlanOp::kIm2col;
  p.src = x; p.dst = y; p.i0 = n_img; p.i1 = H; p.i2 = W; p.i3 = C; p.i4 = stride;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_im2col_s2(sonic_plan_t h, const void* x, void* y, int32_t n_img, int32_t H, int32_t W,
                             int32_t C) {
  SONIC_REQUIRE(h && x && y, "sonic_plan_add_im2col_s2: null argument");
  PlanOp p;
  p.kind = PlanOp::kIm2col;
  p.src = x; p.dst = y; p.i0 = n_img; p.i1 = H; p.i2 = W; p.i3 = C; p.i4 = 2;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_softmax_rows(sonic_plan_t h, void* x, int32_t rows, int32_t cols, int64_t ld, float scale) {
  SONIC_REQUIRE(h && x, "sonic_plan_add_softmax_rows: null argument");
  PlanOp p;
  p.kind = PlanOp::kSoftmaxRows;
  p.dst = x; p.i0 = rows; p.i1 = cols; p.l0 = static_cast<long>(ld); p.eps = scale;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_ln_side(sonic_plan_t h, const float* partials, int32_t parts, int32_t M, int32_t K, float eps,
                           void* side, float* rstd) {
  SONIC_REQUIRE(h && partials && side && rstd, "sonic_plan_add_ln_side: null argument");
  PlanOp p;
  p.kind = PlanOp::kLnSide;
  p.f0 = partials; p.i0 = parts; p.i1 = M; p.i2 = K; p.eps = eps; p.dst = side; p.f1 = rstd;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_timestep_embedding(sonic_plan_t h, const float* t_dev, int32_t dim, float* out) {
  SONIC_REQUIRE(h && t_dev && out && dim % 2 == 0, "sonic_plan_add_timestep_embedding: bad argument");
  PlanOp p;
  p.kind = PlanOp::kTimeEmb;
  p.f0 = t_dev; p.i0 = dim; p.dst = out;
  static_cast<Plan*>(h)->launches += 1;
  static_cast<Plan*>(h)->ops.push_back(p);
  return 0;
}

int sonic_plan_add_gemv(sonic_plan_t h, int32_t n_jobs, const void* const* w, const float* const* bias,
                        const float* const* add, float* const* y, const int32_t* N, const float* x, int32_t K,
                        int32_t silu_in) {
  SONIC_REQUIRE(h && n_jobs > 0 && w && y && N && x, "sonic_plan_add_gemv: bad argument");
  std::vector<GemvJob> jobs(n_jobs);
  int rows = 0;
  for (int j = 0; j < n_jobs; ++j) {
    jobs[j].w = static_cast<const __nv_bfloat16*>(w[j]);
    jobs[j].bias = bias ? bias[j] : nullptr;
    jobs[j].add = add ? add[j] : nullptr;
    jobs[j].y = y[j];
    jobs[j].N = N[j];
    jobs[j].row_start = rows;
    rows += N[j];
  }
  PlanOp p;
  p.kind = PlanOp::kGemv;
  SONIC_CUDA(cudaMalloc(&p.jobs_dev, sizeof(GemvJob) * n_jobs));
  SONIC_CUDA(cudaMemcpy(p.jobs_dev, jobs.data(), sizeof(GemvJob) * n_jobs, cudaMemcpyHostToDevice));
  p.i0 = n_jobs; p.i1 = rows; p.f0 = x; p.i2 = K; p.i3 = silu_in;
  Plan* plan = static_cast<Plan*>(h);
  plan->flops += 2.0 * rows * K;
  plan->launches += 1;
  plan->ops.push_back(p);
  return 0;
}

int sonic_plan_run(sonic_plan_t h, sonic_stream_t stream) {
  SONIC_REQUIRE(h != nullptr, "sonic_plan_run: null plan");
  Plan* plan = static_cast<Plan*>(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (plan->exec) {
    SONIC_CUDA(cudaGraphLaunch(plan->exec, s));
    return 0;
  }
  return run_ops(*plan, s);
}

int sonic_plan_capture(sonic_plan_t h, sonic_stream_t stream) {
  SONIC_REQUIRE(h != nullptr, "sonic_plan_capture: null plan");
  Plan* plan = static_cast<Plan*>(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SONIC_REQUIRE(s != nullptr, "sonic_plan_capture: needs a non-default stream");
  if (plan->exec) { cudaGraphExecDestroy(plan->exec); plan->exec = nullptr; }
  if (plan->graph) { cudaGraphDestroy(plan->graph); plan->graph = nullptr; }
  // one eager run first: sets kernel attributes and surfaces launch errors outside capture
  if (int rc = run_ops(*plan, s)) return rc;
  SONIC_CUDA(cudaStreamSynchronize(s));
  SONIC_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = run_ops(*plan, s);
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (rc) { if (g) cudaGraphDestroy(g); return rc; }
  SONIC_CUDA(e);
  plan->graph = g;
  SONIC_CUDA(cudaGraphInstantiate(&plan->exec, g, 0));
  return 0;
}

// Per-operator device times of one eager (non-graph) run, measured with CUDA events on `stream`.
// kinds[i]: 0 gemm/conv, 1 attention, 2 groupnorm, 3 layernorm, 4 layout/elementwise, 5 gemv/time.
int sonic_plan_profile(sonic_plan_t h, sonic_stream_t stream, int32_t max_ops, float* ms, int32_t* kinds,
                       double* flops, int32_t* n_ops) {
  SONIC_REQUIRE(h && ms && kinds && flops && n_ops, "sonic_plan_profile: null argument");
  Plan* plan = static_cast<Plan*>(h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(plan->ops.size());
  SONIC_REQUIRE(n <= max_ops, "sonic_plan_profile: plan has %d ops, buffer holds %d", n, max_ops);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) SONIC_CUDA(cudaEventCreate(&e));
  Plan one;
  one.ops.resize(1);
  int rc = 0;
  SONIC_CUDA(cudaEventRecord(ev[0], s));
  for (int i = 0; i < n && !rc; ++i) {
    one.ops[0] = plan->ops[i];
    rc = run_ops(one, s);
    cudaEventRecord(ev[i + 1], s);
  }
  for (auto& op : one.ops) { op.att = nullptr; op.jobs_dev = nullptr; }   // borrowed, not owned
  if (!rc) rc = check_cuda(cudaStreamSynchronize(s), "sync", __FILE__, __LINE__);
  for (int i = 0; i < n && !rc; ++i) {
    cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    const PlanOp& op = plan->ops[i];
    flops[i] = 0;
    switch (op.kind) {
      case PlanOp::kGemm: kinds[i] = 0; flops[i] = op.gemm.flops; break;
      case PlanOp::kAttention: kinds[i] = 1; flops[i] = attention_plan_flops(op.att); break;
      case PlanOp::kGroupNorm: kinds[i] = 2; break;
      case PlanOp::kLayerNorm: case PlanOp::kLnSide: kinds[i] = 3; break;
      case PlanOp::kGemv: case PlanOp::kTimeEmb: kinds[i] = 5; break;
      default: kinds[i] = 4; break;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  *n_ops = n;
  return rc;
}

int sonic_plan_stats(sonic_plan_t h, int32_t* n_launches, double* flops) {
  SONIC_REQUIRE(h != nullptr, "sonic_plan_stats: null plan");
  Plan* plan = static_cast<Plan*>(h);
  if (n_launches) *n_launches = plan->launches;
  if (flops) *flops = plan->flops;
  return 0;
}

}  // extern "C"
