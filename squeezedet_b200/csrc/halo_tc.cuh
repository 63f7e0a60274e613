// Halo-tile 3x3 convolution on tcgen05 (SQDET_MATH_TF32X3_TC): the input halo of an output tile is
// fetched ONCE per 16-channel chunk, split hi/lo ONCE, and all nine filter taps read it from shared
// memory through start-address offsets of one SWIZZLE_NONE descriptor.  See halo_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqdet {

struct HaloConvPlan {
  bool enabled = false;
  int B = 0, H = 0, W = 0, Cin = 0, Cout = 0, relu = 1;
  int launches = 1;              // 2 when the conv runs as split-K partials + reduction
  void* impl = nullptr;
};

// 1 = shape taken (plan->enabled), 0 = left to conv_tc.cu, negative = error.
// Replaces ModelSkeleton._conv_layer / _conv_bn_layer for 3x3, stride 1, SAME convolutions
// (src/nn_skeleton.py:374-586): the ConvDet head (src/nets/squeezeDet.py:73-78) and the 3x3 bodies.
int halo_conv_plan(HaloConvPlan* plan, int B, int H, int W, int Cin, int Cout, int relu,
                   bool has_affine, int y_cstride, int y_coff, const float* x_dev, float* y_dev);
int halo_conv_pack_weights(HaloConvPlan* plan, const float* w_hwio, const float* bias);
int halo_conv_set_affine(HaloConvPlan* plan, const float* scale, const float* shift);
int launch_halo_conv(const HaloConvPlan& plan, cudaStream_t stream);
void halo_conv_release(HaloConvPlan* plan);

// split-K reduction shared with conv_tc.cu: y = act((bias + sum_s part[s]) [*scale + shift])
int launch_splitk_reduce(const float* part, float* y, const float* bias, const float* scale,
                         const float* shift, long long npix, int cout, int pitch, int ksplit,
                         int y_cstride, int y_coff, int relu, cudaStream_t stream);

}  // namespace sqdet
