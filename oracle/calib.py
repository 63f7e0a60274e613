"""Single-pass, data-driven rescaling of synthetic weights so that every conv's
pre-activation has unit standard deviation (head: `head_std`) on a synthetic batch
(oracle — test infrastructure only; SURVEY.md §8d "Weights").

conv+bias+ReLU is positively homogeneous, so scaling a layer's kernel and bias by g
scales its output by g exactly: one forward pass suffices.  The per-layer gains are
what `tests/golden/make_calibration.py` commits into
`squeezedet_b200/utils/synth_gains.json` for the product's synthetic initialiser.
"""
from __future__ import annotations

import numpy as np

from . import semantics as S
from .nets import _Tracer, NET_BUILDERS


class _Calibrator(_Tracer):
  def __init__(self, weights, dtype, head_std):
    super().__init__(weights, dtype)
    self.gains = {}
    self.head_std = head_std
    self._last_conv = None

  def _scale(self, scope, z_std, target, leaves):
    g = float(target / max(z_std, 1e-30))
    for leaf in leaves:
      self.w[scope + '/' + leaf] = (self.w[scope + '/' + leaf] * g).astype(np.float32)
    self.gains[scope] = g
    return g

  def conv(self, name, x, filters, size, stride, padding='SAME', relu=True):
    z = S.conv2d(x, self.w[name + '/kernels'], self.w[name + '/biases'], stride, padding,
                 False, dtype=self.dtype)
    target = 1.0 if relu else self.head_std
    g = self._scale(name, z.std(), target, ('kernels', 'biases'))
    z = z * self.dtype(g)
    return self._rec(name, 'conv', S.relu(z) if relu else z)

  def conv_bn(self, scope, x, filters, size, stride, relu=True, bias=False, eps=1e-5):
    # scale gamma and beta: y = bn(conv) is affine in (gamma, beta)
    y = super().conv_bn(scope, x, filters, size, stride, relu=False, bias=bias, eps=eps)
    self.table.pop()
    g = self._scale(scope, y.std(), 1.0, ('gamma', 'beta'))
    y = y * self.dtype(g)
    return self._rec(scope, 'conv_bn', S.relu(y) if relu else y)


def calibrate(net, weights, images, head_std=1.5, n_out=72, dtype=np.float32):
  """Returns (calibrated weights dict, {scope: gain}).  `weights` is not modified."""
  w = {k: np.array(v, dtype=np.float32) for k, v in weights.items()}
  t = _Calibrator(w, dtype, head_std)
  NET_BUILDERS[net](t, np.asarray(images, dtype=dtype), n_out)
  return w, t.gains
