"""Adversarial operand distributions for the tensor-core convolution (conv_tc.cu): the tcgen05
accumulator adds with truncation, and the kernels compensate a segment of m chained MMAs with the
scalar 1 + bias_comp*m measured by tools/mma_bias.cu on one-signed chains.  That model is exact
only when the addends share the sum's sign, so these cases stress the others: all-positive (bias
fully present), heavy-tailed (log-normal: a few products dominate), cancellation-dominated (the sum
is tiny against sum |products|), plus a long K.  Error is measured per output element against the
fp64 oracle on the scale that bounds ANY fp32 summation of the same products,
    |err| <= tol * (|x| (*) |w|),
so a case cannot hide behind a large max.  Both math modes must meet the same bar."""
import numpy as np
import pytest

import oracle
from squeezedet_b200 import _lib
from gpu_util import conv2d_gpu

pytestmark = pytest.mark.gpu

# Forward-error bar in units of sum |products|, for K products per output: fp32 round-to-nearest
# accumulation random-walks to ~ sqrt(K) * 2^-24 (measured 2.7e-6 at K = 2304 on all-positive
# operands with the FFMA kernel); the 3xTF32 path drops terms of 2^-21 per product and carries the
# residual of the scalar bias compensation (tools/mma_bias.cu: a one-signed 36-MMA segment is short
# by 1.6e-6, 1.1e-6 after the compensation).  One bar for both math modes:
def adv_tol(K):
  return 1.2e-7 * np.sqrt(K)


def _case(kind, rng, shape_x, shape_w):
  if kind == 'positive':
    x = rng.uniform(0.5, 1.5, shape_x)
    w = rng.uniform(0.5, 1.5, shape_w)
  elif kind == 'lognormal':
    x = rng.lognormal(0.0, 1.5, shape_x)
    w = rng.lognormal(0.0, 1.5, shape_w) * rng.choice([-1.0, 1.0], shape_w)
  elif kind == 'lognormal_positive':
    x = rng.lognormal(0.0, 1.5, shape_x)
    w = rng.lognormal(0.0, 1.5, shape_w)
  elif kind == 'cancel':
    # per output: +a and -a pairs along the input channels, plus a small residue
    x = np.abs(rng.normal(size=shape_x)) + 0.5
    w = rng.normal(size=shape_w)
    half = shape_w[2] // 2
    w[:, :, half:2 * half, :] = -w[:, :, :half, :]
    x[..., half:2 * half] = x[..., :half] * (1.0 + 1e-3 * rng.normal(size=x[..., :half].shape))
  else:
    raise ValueError(kind)
  return x.astype(np.float32), w.astype(np.float32)


SHAPES = [
    # B, H, W, Cin, Cout, k
    (1, 12, 20, 96, 64, 3),      # 864 products per output, 3 segments of 36 MMAs
    (1, 9, 17, 256, 32, 3),      # 2304 products, K chunks of 32
    (1, 16, 24, 768, 16, 1),     # the longest 1x1 of SqueezeDet (fire11 squeeze)
]


@pytest.mark.parametrize('math_mode', [_lib.MATH_FP32_SIMT, _lib.MATH_TF32X3_TC])
@pytest.mark.parametrize('kind', ['positive', 'lognormal', 'lognormal_positive', 'cancel'])
@pytest.mark.parametrize('shape', SHAPES)
def test_conv_adversarial_operands(shape, kind, math_mode, gpu_device):
  B, H, W, Cin, Cout, k = shape
  rng = np.random.default_rng(1000 + Cin + k)
  x, w = _case(kind, rng, (B, H, W, Cin), (k, k, Cin, Cout))
  want = oracle.conv2d(x, w, None, 1, 'SAME', apply_relu=False, dtype=np.float64)
  bound = oracle.conv2d(np.abs(x), np.abs(w), None, 1, 'SAME', apply_relu=False, dtype=np.float64)
  got = conv2d_gpu(x, w, None, 1, 'SAME', relu=False, math_mode=math_mode)
  ratio = np.abs(got.astype(np.float64) - want) / bound
  assert not np.isnan(got).any()
  tol = adv_tol(k * k * Cin)
  assert ratio.max() < tol, (kind, shape, float(ratio.max()), tol)
  # the error must not be mostly one-sided: the MEAN signed error stays below half the bar
  signed = ((got.astype(np.float64) - want) / bound).mean()
  assert abs(signed) < tol / 2, (kind, shape, float(signed), tol)
