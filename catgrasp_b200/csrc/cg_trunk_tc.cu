// cg_trunk_tc.cu -- tcgen05 trunk (engine 1). Placeholder until the tensor-core kernel lands.
#include "cg_net.cuh"
size_t cg_tc_w3_bytes() { return 256; }
int cg_tc_prepare_w3(cg_ctx *, const float *, void *) { return CG_OK; }
int cg_trunk_launch_tc(cg_ctx *ctx, const cg_trunk_args &) {
  ctx->err = "engine 1 (tcgen05) not built";
  return CG_EUNSUPPORTED;
}
