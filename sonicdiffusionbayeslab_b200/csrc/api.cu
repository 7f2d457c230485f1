// extern "C" entry points of libsonic (declared in include/sonic.h).
#include "../../include/sonic.h"

#include "common.cuh"
#include "gemm.cuh"

using namespace sonic;

extern "C" {

const char* sonic_last_error(void) { return last_error(); }
const char* sonic_version(void) { return "sonic-b200 0.1 sm_100a"; }

int sonic_conv_gemm(const sonic_gemm_args* a, sonic_stream_t stream) {
  SONIC_REQUIRE(a != nullptr, "sonic_conv_gemm: null args");
  GemmOp op;
  op.a0 = a->a0; op.c0 = a->c0; op.ld0 = a->ld0;
  op.a1 = a->a1; op.c1 = a->c1; op.ld1 = a->ld1;
  op.n_img = a->n_img; op.H = a->H; op.W = a->W;
  op.w = a->w; op.N = a->N; op.taps = a->taps;
  op.bias = a->bias; op.row_bias = a->row_bias;
  op.residual = a->residual; op.ld_res = a->ld_res;
  op.out = a->out; op.ld_out = a->ld_out;
  op.epilogue = a->epilogue; op.block_n = a->block_n;
  GemmPlan plan;
  if (int rc = gemm_plan(op, &plan)) return rc;
  return gemm_launch(plan, static_cast<cudaStream_t>(stream));
}

int sonic_gemm_block_n(int32_t N, int32_t n_img, int32_t H, int32_t W, int32_t epilogue) {
  return gemm_choose_block_n(N, n_img, H, W, epilogue);
}

}  // extern "C"
