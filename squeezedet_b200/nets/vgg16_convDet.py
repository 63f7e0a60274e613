"""VGG16+ConvDet — drop-in for reference ``src/nets/vgg16_convDet.py``: the 13 VGG16
3x3 convs in five blocks (variables scoped ``convN/convN_i``), 2x2/2 SAME pools
after blocks 1-4, then the ConvDet head ``conv6``."""
from __future__ import annotations

from ..nn_skeleton import ModelSkeleton

# (block, number of 3x3 convs, channels) — reference vgg16_convDet.py:41-83
_VGG_BLOCKS = ((1, 2, 64), (2, 2, 128), (3, 3, 256), (4, 3, 512), (5, 3, 512))


class VGG16ConvDet(ModelSkeleton):
  def __init__(self, mc, gpu_id=0, math_mode=None):
    ModelSkeleton.__init__(self, mc, gpu_id, math_mode)
    self._add_forward_graph()
    self._add_interpretation_graph()

  def _add_forward_graph(self):
    mc = self.mc
    x = self.image_input
    for block, count, channels in _VGG_BLOCKS:
      for i in range(1, count + 1):
        # tf.variable_scope('convN') + layer 'convN_i' -> 'convN/convN_i/kernels'
        x = self._conv_layer('conv%d/conv%d_%d' % (block, block, i), x,
                             filters=channels, size=3, stride=1,
                             freeze=block <= 2)
      if block < 5:
        x = self._pooling_layer('pool%d' % block, x, size=2, stride=2)
    x = self._dropout(x, self.keep_prob, name='drop6')
    num_output = mc.ANCHOR_PER_GRID * (mc.CLASSES + 1 + 4)
    self.preds = self._conv_layer('conv6', x, filters=num_output, size=3, stride=1,
                                  padding='SAME', xavier=False, relu=False,
                                  stddev=0.0001)
