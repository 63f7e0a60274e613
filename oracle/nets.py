"""The four reference topologies restated over ``oracle.semantics``
(oracle — test infrastructure only).

Follows (reference, read-only):
  * ``src/nets/squeezeDet.py:30-106``       SqueezeDet      (conv1 3x3/2 SAME, SAME pools)
  * ``src/nets/squeezeDetPlus.py:30-106``   SqueezeDet+     (conv1 7x7/2 VALID, VALID pools)
  * ``src/nets/vgg16_convDet.py:31-90``     VGG16+ConvDet   (2x2/2 SAME pools)
  * ``src/nets/resnet50_convDet.py:31-169`` ResNet50+ConvDet (frozen BN, conv1..conv4_x)
Weights are addressed by the reference's TF variable names
(``<scope>/kernels`` HWIO, ``<scope>/biases``; BN ``gamma/beta/mean/var``),
which is what ``tf.train.Saver(model.model_params)`` stores (demo.py:181).
Dropout is the identity at inference (nn_skeleton.py:78).
"""
from __future__ import annotations

import numpy as np

from . import semantics as S


class _Tracer:
  """Runs ops and (optionally) records per-layer shapes/FLOPs/outputs."""

  def __init__(self, weights, dtype, keep=None):
    self.w = weights
    self.dtype = dtype
    self.table = []          # (name, kind, out_shape[1:], 2*MAC flops, params)
    self.keep = keep         # None or dict to be filled with named outputs

  def _rec(self, name, kind, y, flops=0, params=0):
    self.table.append((name, kind, tuple(y.shape[1:]), int(flops), int(params)))
    if self.keep is not None:
      self.keep[name] = y
    return y

  def conv(self, name, x, filters, size, stride, padding='SAME', relu=True):
    w = self.w[name + '/kernels']
    b = self.w[name + '/biases']
    assert w.shape == (size, size, x.shape[3], filters), (name, w.shape)
    y = S.conv2d(x, w, b, stride, padding, relu, dtype=self.dtype)
    fl = 2 * size * size * x.shape[3] * filters * y.shape[1] * y.shape[2]
    return self._rec(name, 'conv', y, fl, (1 + size * size * x.shape[3]) * filters)

  def conv_bn(self, scope, x, filters, size, stride, relu=True, bias=False,
              eps=1e-5):
    w = self.w[scope + '/kernels']
    assert w.shape == (size, size, x.shape[3], filters), (scope, w.shape)
    b = self.w[scope + '/biases'] if bias else None
    y = S.conv2d(x, w, b, stride, 'SAME', False, dtype=self.dtype)
    y = S.batch_norm_frozen(y, self.w[scope + '/mean'], self.w[scope + '/var'],
                            self.w[scope + '/beta'], self.w[scope + '/gamma'],
                            eps)
    if relu:
      y = S.relu(y)
    fl = 2 * size * size * x.shape[3] * filters * y.shape[1] * y.shape[2]
    return self._rec(scope, 'conv_bn', y.astype(self.dtype), fl,
                     (1 + size * size * x.shape[3]) * filters)

  def pool(self, name, x, size, stride, padding='SAME'):
    return self._rec(name, 'pool', S.max_pool(x, size, stride, padding))

  def fire(self, name, x, s1x1, e1x1, e3x3):
    """squeezeDet.py:81-106: concat_C(relu(1x1_e1(q)), relu(3x3_e3(q))),
    q = relu(1x1_s(x))."""
    q = self.conv(name + '/squeeze1x1', x, s1x1, 1, 1)
    a = self.conv(name + '/expand1x1', q, e1x1, 1, 1)
    b = self.conv(name + '/expand3x3', q, e3x3, 3, 1)
    return self._rec(name, 'concat', np.concatenate([a, b], axis=3))


def _squeezedet(t, x, n_out):
  x = t.conv('conv1', x, 64, 3, 2, 'SAME')
  x = t.pool('pool1', x, 3, 2, 'SAME')
  x = t.fire('fire2', x, 16, 64, 64)
  x = t.fire('fire3', x, 16, 64, 64)
  x = t.pool('pool3', x, 3, 2, 'SAME')
  x = t.fire('fire4', x, 32, 128, 128)
  x = t.fire('fire5', x, 32, 128, 128)
  x = t.pool('pool5', x, 3, 2, 'SAME')
  x = t.fire('fire6', x, 48, 192, 192)
  x = t.fire('fire7', x, 48, 192, 192)
  x = t.fire('fire8', x, 64, 256, 256)
  x = t.fire('fire9', x, 64, 256, 256)
  x = t.fire('fire10', x, 96, 384, 384)
  x = t.fire('fire11', x, 96, 384, 384)
  return t.conv('conv12', x, n_out, 3, 1, 'SAME', relu=False)


def _squeezedet_plus(t, x, n_out):
  x = t.conv('conv1', x, 96, 7, 2, 'VALID')
  x = t.pool('pool1', x, 3, 2, 'VALID')
  x = t.fire('fire2', x, 96, 64, 64)
  x = t.fire('fire3', x, 96, 64, 64)
  x = t.fire('fire4', x, 192, 128, 128)
  x = t.pool('pool4', x, 3, 2, 'VALID')
  x = t.fire('fire5', x, 192, 128, 128)
  x = t.fire('fire6', x, 288, 192, 192)
  x = t.fire('fire7', x, 288, 192, 192)
  x = t.fire('fire8', x, 384, 256, 256)
  x = t.pool('pool8', x, 3, 2, 'VALID')
  x = t.fire('fire9', x, 384, 256, 256)
  x = t.fire('fire10', x, 384, 256, 256)
  x = t.fire('fire11', x, 384, 256, 256)
  return t.conv('conv12', x, n_out, 3, 1, 'SAME', relu=False)


def _vgg16(t, x, n_out):
  cfg = [(1, 2, 64), (2, 2, 128), (3, 3, 256), (4, 3, 512), (5, 3, 512)]
  for blk, n, ch in cfg:
    for i in range(1, n + 1):
      x = t.conv('conv%d/conv%d_%d' % (blk, blk, i), x, ch, 3, 1, 'SAME')
    if blk < 5:
      x = t.pool('pool%d' % blk, x, 2, 2, 'SAME')
  return t.conv('conv6', x, n_out, 3, 1, 'SAME', relu=False)


def _resnet50(t, x, n_out):
  x = t.conv_bn('conv1', x, 64, 7, 2, relu=True, bias=True)
  x = t.pool('pool1', x, 3, 2, 'VALID')
  stages = [('2', 'abc', 64, 256, False), ('3', 'abcd', 128, 512, True),
            ('4', 'abcdef', 256, 1024, True)]
  for sid, blocks, mid, out, down in stages:
    for blk in blocks:
      unit = sid + blk
      scope = 'conv%s_x/res%s/' % (sid, unit)
      first = blk == 'a'
      stride = 2 if (first and down) else 1
      b2 = scope + 'res%s_branch2/res%s_branch2' % (unit, unit)
      if first:     # branch1 is built first in the reference (resnet50_convDet.py:49-54)
        sc = t.conv_bn(scope + 'res%s_branch1' % unit, x, out, 1, stride,
                       relu=False)
      else:
        sc = x
      y = t.conv_bn(b2 + 'a', x, mid, 1, stride, relu=True)
      y = t.conv_bn(b2 + 'b', y, mid, 3, 1, relu=True)
      y = t.conv_bn(b2 + 'c', y, out, 1, 1, relu=False)
      x = t._rec('res' + unit, 'add_relu', S.relu(sc + y))
  return t.conv('conv5', x, n_out, 3, 1, 'SAME', relu=False)


NET_BUILDERS = {
    'squeezeDet': _squeezedet,
    'squeezeDet+': _squeezedet_plus,
    'vgg16': _vgg16,
    'resnet50': _resnet50,
}


def forward(net, weights, images, n_out=72, dtype=np.float32, keep=None):
  """images [B,H,W,3] (BGR, mean-subtracted) -> preds [B,Hg,Wg,n_out].
  `keep`: optional dict filled with every named intermediate tensor."""
  t = _Tracer(weights, dtype, keep)
  x = np.asarray(images, dtype=dtype)
  return NET_BUILDERS[net](t, x, n_out)


def layer_table(net, height, width, n_out=72, specs=None):
  """(name, kind, out_shape, flops, params) per layer on a 1-pixel-cheap trace:
  geometry is computed analytically (no arithmetic on real-size tensors).
  `specs`: optional list filled with (param_name, shape) in reference order."""
  rows = []
  if specs is None:
    specs = []

  class T:
    def conv(self, name, x, filters, size, stride, padding='SAME', relu=True):
      h, w, c = x
      if not getattr(self, '_bn', False):
        specs.append((name + '/kernels', (size, size, c, filters)))
        specs.append((name + '/biases', (filters,)))
      ho = S.conv_geometry(h, size, stride, padding)[0]
      wo = S.conv_geometry(w, size, stride, padding)[0]
      rows.append((name, 'conv', (ho, wo, filters),
                   2 * size * size * c * filters * ho * wo,
                   (1 + size * size * c) * filters))
      return (ho, wo, filters)

    def conv_bn(self, scope, x, filters, size, stride, relu=True, bias=False,
                eps=1e-5):
      specs.append((scope + '/kernels', (size, size, x[2], filters)))
      for leaf in (['biases'] if bias else []) + ['gamma', 'beta', 'mean', 'var']:
        specs.append((scope + '/' + leaf, (filters,)))
      self._bn = True
      try:
        return self.conv(scope, x, filters, size, stride, 'SAME', relu)
      finally:
        self._bn = False

    def pool(self, name, x, size, stride, padding='SAME'):
      h, w, c = x
      y = (S.conv_geometry(h, size, stride, padding)[0],
           S.conv_geometry(w, size, stride, padding)[0], c)
      rows.append((name, 'pool', y, 0, 0))
      return y

    def fire(self, name, x, s1x1, e1x1, e3x3):
      q = self.conv(name + '/squeeze1x1', x, s1x1, 1, 1)
      a = self.conv(name + '/expand1x1', q, e1x1, 1, 1)
      b = self.conv(name + '/expand3x3', q, e3x3, 3, 1)
      y = (a[0], a[1], a[2] + b[2])
      rows.append((name, 'concat', y, 0, 0))
      return y

    def _rec(self, name, kind, y, flops=0, params=0):
      rows.append((name, kind, y, 0, 0))
      return y

  class _Sym(tuple):
    """(h, w, c) that tolerates `sc + y` and relu() in the resnet builder."""
    def __add__(self, other):
      assert tuple(self) == tuple(other), (self, other)
      return self

  t = T()
  orig_relu = S.relu
  S.relu = lambda v: v
  try:
    # wrap results so residual adds work on shape tuples
    for meth in ('conv', 'conv_bn', 'pool', 'fire', '_rec'):
      fn = getattr(t, meth)
      setattr(t, meth, (lambda f: (lambda *a, **k: _Sym(f(*a, **k))))(fn))
    NET_BUILDERS[net](t, _Sym((height, width, 3)), n_out)
  finally:
    S.relu = orig_relu
  return rows


def param_specs(net, n_out=72):
  """[(reference variable name, shape)] in the order of model.model_params."""
  specs = []
  layer_table(net, 375, 1242, n_out, specs)
  return specs
