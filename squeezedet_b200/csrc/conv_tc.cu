// placeholder — replaced by the tcgen05 implementation
#include "common.cuh"
#include "conv_tc.cuh"
namespace sqdet {
int tc_conv_plan(TcConvPlan* plan, int, int, int, int, int, int, int, int, int, bool, int, int,
                 const float*, float*) { plan->enabled = false; return 0; }
int tc_fire_plan(TcFirePlan* plan, int, int, int, int, int, int, const float*, float*) {
  plan->enabled = false; return 0; }
int tc_conv_pack_weights(TcConvPlan*, const float*, const float*) { return 0; }
int tc_fire_pack_weights(TcFirePlan*, const float*, const float*, const float*, const float*) { return 0; }
int launch_conv_tc(const TcConvPlan&, const float*, float*, cudaStream_t) {
  return fail(SQDET_ERR_UNSUPPORTED, "tensor-core conv not built"); }
int launch_fire_expand_tc(const TcFirePlan&, const float*, float*, cudaStream_t) {
  return fail(SQDET_ERR_UNSUPPORTED, "tensor-core fire not built"); }
void tc_conv_release(TcConvPlan*) {}
void tc_fire_release(TcFirePlan*) {}
int conv2d_tc_oneshot(const float*, const float*, const float*, const float*, const float*, float*,
                      int, int, int, int, int, int, int, int, int, int, int, cudaStream_t) {
  return fail(SQDET_ERR_UNSUPPORTED, "tensor-core conv not built"); }
}
