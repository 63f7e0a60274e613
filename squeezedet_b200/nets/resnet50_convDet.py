"""ResNet50+ConvDet — drop-in for reference ``src/nets/resnet50_convDet.py``:
conv1 7x7/2 (+bias, frozen BN) -> pool1 3x3/2 VALID -> conv2_x (3 units) ->
conv3_x (4, stride 2) -> conv4_x (6, stride 2) -> ConvDet head ``conv5``.
BatchNorm uses stored statistics (an affine per channel, folded into the conv
epilogue); every residual unit ends in relu(shortcut + branch2)."""
from __future__ import annotations

from ..nn_skeleton import ModelSkeleton

# (stage id, unit letters, bottleneck width, output width, down-sample first unit)
_STAGES = (('2', 'abc', 64, 256, False),
           ('3', 'abcd', 128, 512, True),
           ('4', 'abcdef', 256, 1024, True))


class ResNet50ConvDet(ModelSkeleton):
  def __init__(self, mc, gpu_id=0, math_mode=None):
    ModelSkeleton.__init__(self, mc, gpu_id, math_mode)
    self._add_forward_graph()
    self._add_interpretation_graph()

  def _res_branch(self, inputs, layer_name, in_filters, out_filters,
                  down_sample=False, freeze=False, scope=''):
    """Residual branch: 1x1(/2) -> 3x3 -> 1x1, BN after each, relu after the first
    two (reference resnet50_convDet.py:134-169).  `scope` is the enclosing TF
    variable scope ('conv2_x/res2a/'), so names match the reference checkpoint."""
    base = '%sres%s_branch2/' % (scope, layer_name)
    stride = 2 if down_sample else 1
    plan = (('a', in_filters, 1, stride, True), ('b', in_filters, 3, 1, True),
            ('c', out_filters, 1, 1, False))
    out = inputs
    for suffix, filters, size, strd, relu in plan:
      out = self._conv_bn_layer(
          out, conv_param_name='%sres%s_branch2%s' % (base, layer_name, suffix),
          bn_param_name='bn%s_branch2%s' % (layer_name, suffix),
          scale_param_name='scale%s_branch2%s' % (layer_name, suffix),
          filters=filters, size=size, stride=strd, freeze=freeze, relu=relu)
    return out

  def _add_forward_graph(self):
    mc = self.mc
    x = self._conv_bn_layer(self.image_input, 'conv1', 'bn_conv1', 'scale_conv1',
                            filters=64, size=7, stride=2, freeze=True,
                            conv_with_bias=True)
    x = self._pooling_layer('pool1', x, size=3, stride=2, padding='VALID')
    for sid, letters, mid, width, down in _STAGES:
      for letter in letters:
        unit = sid + letter
        scope = 'conv%s_x/res%s/' % (sid, unit)
        first = letter == 'a'
        if first:   # projection shortcut (1x1, stride 2 when down-sampling), no relu
          shortcut = self._conv_bn_layer(
              x, '%sres%s_branch1' % (scope, unit), 'bn%s_branch1' % unit,
              'scale%s_branch1' % unit, filters=width, size=1,
              stride=2 if down else 1, freeze=sid != '4', relu=False)
        else:
          shortcut = x
        branch2 = self._res_branch(x, layer_name=unit, in_filters=mid,
                                   out_filters=width, down_sample=first and down,
                                   freeze=sid != '4', scope=scope)
        x = self._add_relu('res' + unit, shortcut, branch2)
    x = self._dropout(x, self.keep_prob, name='drop4')
    num_output = mc.ANCHOR_PER_GRID * (mc.CLASSES + 1 + 4)
    self.preds = self._conv_layer('conv5', x, filters=num_output, size=3, stride=1,
                                  padding='SAME', xavier=False, relu=False,
                                  stddev=0.0001)
