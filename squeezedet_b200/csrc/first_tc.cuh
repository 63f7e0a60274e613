// First layer on tcgen05: conv (3-channel image, KxK, stride 2) + bias [+ frozen BN] + ReLU +
// stride-2 max-pool in ONE kernel, pooled-pixel-major ("the pool is a max over accumulators").
// See first_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqdet {

struct FirstTcPlan {
  bool enabled = false;
  int B = 0, H = 0, W = 0, Cout = 0, ksize = 3;
  void* impl = nullptr;
};

// 1 = shape taken (plan->enabled), 0 = left to the FFMA kernel, negative = error.
// Replaces _conv_layer / _conv_bn_layer + _pooling_layer of the first layer
// (src/nets/squeezeDet.py:40-44, squeezeDetPlus.py:40-44, resnet50_convDet.py:41-48).
int first_tc_plan(FirstTcPlan* plan, int B, int H, int W, int Cout, int ksize, int stride,
                  int conv_padding, int relu, bool has_affine, int pool_size, int pool_stride,
                  int pool_padding, float* y_dev);
int first_tc_pack_weights(FirstTcPlan* plan, const float* w_hwio, const float* bias);
int first_tc_set_affine(FirstTcPlan* plan, const float* scale, const float* shift);
int launch_first_tc(const FirstTcPlan& plan, const float* x_dev, cudaStream_t stream);
void first_tc_release(FirstTcPlan* plan);

}  // namespace sqdet
