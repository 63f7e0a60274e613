"""torch-CPU backend of the oracle's forward pass (oracle — test infrastructure only).

Same semantics as ``oracle.semantics`` (explicit asymmetric TF-SAME padding, then a VALID
conv / -inf padded max-pool), but through torch's multi-threaded CPU kernels: this is the
"CPU restatement (not TF)" that bench.py times on the GPU box's host cores as the
``cpu_baseline`` / ``--impl reference`` arm.  Validated against the
numpy oracle in tests/test_oracle_semantics.py.
"""
from __future__ import annotations

import numpy as np

from . import semantics as S
from .nets import _Tracer, NET_BUILDERS


class _TorchTracer(_Tracer):
  """Activations are NCHW torch tensors internally; weights are cached as OIHW."""

  def __init__(self, weights, dtype):
    import torch
    super().__init__(weights, dtype)
    self.torch = torch
    self.tdt = torch.float32 if dtype == np.float32 else torch.float64
    self._wcache = {}

  def _t(self, name):
    if name not in self._wcache:
      self._wcache[name] = self.torch.from_numpy(np.asarray(self.w[name])).to(self.tdt)
    return self._wcache[name]

  def _kernel(self, name):
    key = name + '#oihw'
    if key not in self._wcache:
      self._wcache[key] = self._t(name).permute(3, 2, 0, 1).contiguous()
    return self._wcache[key]

  def _conv(self, scope, x, size, stride, padding, bias):
    F = self.torch.nn.functional
    H, W = x.shape[2], x.shape[3]
    _, pt, pb = S.conv_geometry(H, size, stride, padding)
    _, pl, pr = S.conv_geometry(W, size, stride, padding)
    if pt or pb or pl or pr:
      x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, self._kernel(scope + '/kernels'),
                    self._t(scope + '/biases') if bias else None, stride)

  def conv(self, name, x, filters, size, stride, padding='SAME', relu=True):
    y = self._conv(name, x, size, stride, padding, True)
    return self.torch.relu_(y) if relu else y

  def conv_bn(self, scope, x, filters, size, stride, relu=True, bias=False, eps=1e-5):
    y = self._conv(scope, x, size, stride, 'SAME', bias)
    inv = self.torch.rsqrt(self._t(scope + '/var') + eps) * self._t(scope + '/gamma')
    shift = self._t(scope + '/beta') - self._t(scope + '/mean') * inv
    y = y * inv.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    return self.torch.relu_(y) if relu else y

  def pool(self, name, x, size, stride, padding='SAME'):
    F = self.torch.nn.functional
    H, W = x.shape[2], x.shape[3]
    _, pt, pb = S.conv_geometry(H, size, stride, padding)
    _, pl, pr = S.conv_geometry(W, size, stride, padding)
    if pt or pb or pl or pr:
      x = F.pad(x, (pl, pr, pt, pb), value=float('-inf'))
    return F.max_pool2d(x, size, stride)

  def fire(self, name, x, s1x1, e1x1, e3x3):
    q = self.conv(name + '/squeeze1x1', x, s1x1, 1, 1)
    a = self.conv(name + '/expand1x1', q, e1x1, 1, 1)
    b = self.conv(name + '/expand3x3', q, e3x3, 3, 1)
    return self.torch.cat([a, b], dim=1)

  def _rec(self, name, kind, y, flops=0, params=0):
    return y


class TorchForward:
  """Reusable forward (weights converted once)."""

  def __init__(self, net, weights, n_out=72, dtype=np.float32, threads=None):
    import torch
    if threads:
      torch.set_num_threads(int(threads))
    self.net, self.n_out = net, n_out
    self.t = _TorchTracer(weights, dtype)
    self._relu = S.relu

  def __call__(self, images):
    torch = self.t.torch
    x = torch.from_numpy(np.ascontiguousarray(images)).to(self.t.tdt).permute(0, 3, 1, 2)
    saved = S.relu
    S.relu = torch.relu          # the resnet builder calls S.relu(shortcut + branch)
    try:
      with torch.no_grad():
        y = NET_BUILDERS[self.net](self.t, x, self.n_out)
    finally:
      S.relu = saved
    return y.permute(0, 2, 3, 1).contiguous().numpy()
