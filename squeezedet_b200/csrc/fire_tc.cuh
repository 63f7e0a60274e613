// Single-kernel fire module on tcgen05 (squeeze 1x1 -> expand 1x1 || 3x3 -> concat), and the
// halo-tile 3x3 convolution that shares its machinery.  See fire_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqdet {

struct FusedFirePlan {
  bool enabled = false;
  int B = 0, H = 0, W = 0, Cin = 0, S = 0, E1 = 0, E3 = 0;
  void* impl = nullptr;
};

// 1 = shape taken (plan->enabled), 0 = left to the two-launch path, negative = error.
int fused_fire_plan(FusedFirePlan* plan, int B, int H, int W, int Cin, int S, int E1, int E3,
                    const float* x_dev, float* y_dev);
int fused_fire_pack_weights(FusedFirePlan* plan, const float* w_sq, const float* b_sq,
                            const float* w_e1, const float* b_e1, const float* w_e3,
                            const float* b_e3);
int launch_fused_fire(const FusedFirePlan& plan, cudaStream_t stream);
void fused_fire_release(FusedFirePlan* plan);

// Stage-isolated entry behind sqdet_fire: device weights in HWIO; returns 0 when it ran,
// 1 when the shape is not taken by the fused kernel, negative on error.  Synchronises.
int fire_fused_oneshot(const float* x_dev, const float* w_sq_dev, const float* b_sq_dev,
                       const float* w_e1_dev, const float* b_e1_dev, const float* w_e3_dev,
                       const float* b_e3_dev, float* y_dev, int B, int H, int W, int Cin, int S,
                       int E1, int E3, cudaStream_t stream);

}  // namespace sqdet
