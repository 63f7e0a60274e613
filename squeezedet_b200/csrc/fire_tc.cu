// placeholder: replaced by the fused kernel
#include "common.cuh"
#include "fire_tc.cuh"
namespace sqdet {
int fused_fire_plan(FusedFirePlan* plan, int, int, int, int, int, int, int, const float*, float*) {
  plan->enabled = false;
  return 0;
}
int fused_fire_pack_weights(FusedFirePlan*, const float*, const float*, const float*, const float*,
                            const float*, const float*) { return SQDET_OK; }
int launch_fused_fire(const FusedFirePlan&, cudaStream_t) { return fail(SQDET_ERR_STATE, "no fused fire plan"); }
void fused_fire_release(FusedFirePlan* plan) { plan->enabled = false; plan->impl = nullptr; }
int fire_fused_oneshot(const float*, const float*, const float*, const float*, const float*,
                       const float*, const float*, float*, int, int, int, int, int, int, int,
                       cudaStream_t) { return 1; }
}  // namespace sqdet
