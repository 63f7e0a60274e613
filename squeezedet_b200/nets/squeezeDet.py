"""SqueezeDet — drop-in for reference ``src/nets/squeezeDet.py`` (class
``SqueezeDet(mc, gpu_id)``): conv1 3x3/2 SAME -> pool1 -> fire2,3 -> pool3 ->
fire4,5 -> pool5 -> fire6..11 -> (dropout = identity) -> conv12 3x3 ConvDet head.
The topology is a table; every row becomes one op of the engine plan."""
from __future__ import annotations

from ..nn_skeleton import ModelSkeleton

# (kind, name, *args) — reference src/nets/squeezeDet.py:40-79
_SQUEEZEDET_BODY = (
    ('conv', 'conv1', 64, 3, 2, 'SAME'),
    ('pool', 'pool1', 3, 2, 'SAME'),
    ('fire', 'fire2', 16, 64, 64),
    ('fire', 'fire3', 16, 64, 64),
    ('pool', 'pool3', 3, 2, 'SAME'),
    ('fire', 'fire4', 32, 128, 128),
    ('fire', 'fire5', 32, 128, 128),
    ('pool', 'pool5', 3, 2, 'SAME'),
    ('fire', 'fire6', 48, 192, 192),
    ('fire', 'fire7', 48, 192, 192),
    ('fire', 'fire8', 64, 256, 256),
    ('fire', 'fire9', 64, 256, 256),
    ('fire', 'fire10', 96, 384, 384),
    ('fire', 'fire11', 96, 384, 384),
)


class FireNetBase(ModelSkeleton):
  """Shared by SqueezeDet and SqueezeDet+: a conv/pool/fire table + ConvDet head."""

  BODY = ()
  HEAD_NAME = 'conv12'

  def __init__(self, mc, gpu_id=0, math_mode=None):
    ModelSkeleton.__init__(self, mc, gpu_id, math_mode)
    self._add_forward_graph()
    self._add_interpretation_graph()
    # _add_loss_graph / _add_train_graph / _add_viz_graph: training-only, not built.

  def _fire_layer(self, layer_name, inputs, s1x1, e1x1, e3x3, stddev=0.01,
                  freeze=False):
    """Fire layer constructor (reference squeezeDet.py:81-106): squeeze 1x1 ->
    {expand 1x1 || expand 3x3} -> channel concat, all with bias + ReLU."""
    return self._fused_fire(layer_name, inputs, s1x1, e1x1, e3x3)

  def _add_forward_graph(self):
    mc = self.mc
    x = self.image_input
    for row in self.BODY:
      kind, name, args = row[0], row[1], row[2:]
      if kind == 'conv':
        filters, size, stride, padding = args
        x = self._conv_layer(name, x, filters=filters, size=size, stride=stride,
                             padding=padding, freeze=True)
      elif kind == 'pool':
        size, stride, padding = args
        x = self._pooling_layer(name, x, size=size, stride=stride, padding=padding)
      else:
        x = self._fire_layer(name, x, s1x1=args[0], e1x1=args[1], e3x3=args[2])
    x = self._dropout(x, self.keep_prob, name='drop11')
    num_output = mc.ANCHOR_PER_GRID * (mc.CLASSES + 1 + 4)
    self.preds = self._conv_layer(self.HEAD_NAME, x, filters=num_output, size=3,
                                  stride=1, padding='SAME', xavier=False,
                                  relu=False, stddev=0.0001)


class SqueezeDet(FireNetBase):
  BODY = _SQUEEZEDET_BODY
