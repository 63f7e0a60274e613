"""Model configuration objects (`mc`) — the constructor argument of every net.

Same keys, values and factory names as the reference's EasyDict configs
(reference src/config/config.py:10-142 ``base_model_config`` and
src/config/kitti_{squeezeDet,squeezeDetPlus,vgg16,res50}_config.py), rebuilt as
one table-driven module: a reference user's ``mc.IMAGE_WIDTH``,
``mc.ANCHOR_BOX``, ``mc.TOP_N_DETECTION`` ... read the same here.  Training-only
keys are kept so reference scripts that touch them do not break; this engine
never reads them.
"""
from __future__ import annotations

import numpy as np


class ModelConfig(dict):
  """Attribute-style dict (stand-in for easydict.EasyDict, which the reference
  imports at src/config/config.py:8 and which is not installed here)."""

  def __getattr__(self, key):
    try:
      return self[key]
    except KeyError as exc:
      raise AttributeError(key) from exc

  def __setattr__(self, key, value):
    self[key] = value

  def copy(self):
    return ModelConfig(dict.copy(self))


_CLASS_NAMES = {
    'PASCAL_VOC': ('aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car',
                   'cat', 'chair', 'cow', 'diningtable', 'dog', 'horse',
                   'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train',
                   'tvmonitor'),
    'KITTI': ('car', 'pedestrian', 'cyclist'),
}

# defaults of base_model_config (reference src/config/config.py:32-140)
_BASE = dict(
    GRID_POOL_WIDTH=7, GRID_POOL_HEIGHT=7, LEAKY_COEF=0.1, KEEP_PROB=0.5,
    IMAGE_WIDTH=224, IMAGE_HEIGHT=224, ANCHOR_PER_GRID=-1, BATCH_SIZE=20,
    PROB_THRESH=0.005, PLOT_PROB_THRESH=0.5, NMS_THRESH=0.2,
    LOSS_COEF_CONF=1.0, LOSS_COEF_CLASS=1.0, LOSS_COEF_BBOX=10.0,
    DECAY_STEPS=10000, LR_DECAY_FACTOR=0.1, LEARNING_RATE=0.005, MOMENTUM=0.9,
    WEIGHT_DECAY=0.0005, LOAD_PRETRAINED_MODEL=True, PRETRAINED_MODEL_PATH='',
    DEBUG_MODE=False, EPSILON=1e-16, EXP_THRESH=1.0, MAX_GRAD_NORM=10.0,
    DATA_AUGMENTATION=False, DRIFT_X=0, DRIFT_Y=0, EXCLUDE_HARD_EXAMPLES=True,
    BATCH_NORM_EPSILON=1e-5, NUM_THREAD=4, QUEUE_CAPACITY=100, IS_TRAINING=False,
)


def base_model_config(dataset='PASCAL_VOC'):
  name = dataset.upper()
  assert name in _CLASS_NAMES, \
      'Currently only support PASCAL_VOC or KITTI dataset'
  mc = ModelConfig(_BASE)
  mc.DATASET = name
  mc.CLASS_NAMES = _CLASS_NAMES[name]
  mc.CLASSES = len(mc.CLASS_NAMES)
  mc.ANCHOR_BOX = []
  mc.ANCHORS = 0
  # BGR mean of VGG16, shape (1,1,3)  (config.py:70-72)
  mc.BGR_MEANS = np.array([[[103.939, 116.779, 123.68]]])
  return mc


# anchor shapes (w, h) per grid cell
_SHAPES_SQUEEZE = ((36., 37.), (366., 174.), (115., 59.), (162., 87.), (38., 90.),
                   (258., 173.), (224., 108.), (78., 170.), (72., 43.))
_SHAPES_RES50 = ((94., 49.), (225., 161.), (170., 91.), (390., 181.), (41., 32.),
                 (128., 64.), (298., 164.), (232., 99.), (65., 42.))


def make_anchor_box(image_width, image_height, grid_h, grid_w, shapes):
  """``set_anchors`` (reference kitti_squeezeDet_config.py:45-79): [A,4] float64
  (cx, cy, w, h) in (row, col, shape) order; centres at (j+1)*W/(grid_w+1),
  (i+1)*H/(grid_h+1) — multiply first, then divide, bit-for-bit as the reference."""
  shapes = np.asarray(shapes, dtype=np.float64)
  col = np.arange(1, grid_w + 1) * float(image_width) / (grid_w + 1)
  row = np.arange(1, grid_h + 1) * float(image_height) / (grid_h + 1)
  box = np.empty((grid_h, grid_w, len(shapes), 4), dtype=np.float64)
  box[:, :, :, 0] = col.reshape(1, grid_w, 1)
  box[:, :, :, 1] = row.reshape(grid_h, 1, 1)
  box[:, :, :, 2:] = shapes.reshape(1, 1, len(shapes), 2)
  return box.reshape(grid_h * grid_w * len(shapes), 4)


# per-net overrides: (IMAGE_WIDTH, IMAGE_HEIGHT, BATCH_SIZE, grid_h, grid_w, shapes)
_KITTI_NETS = {
    'squeezeDet': (1248, 384, 20, 24, 78, _SHAPES_SQUEEZE),
    'squeezeDet+': (1242, 375, 20, 22, 76, _SHAPES_SQUEEZE),
    'vgg16': (1242, 375, 5, 24, 78, _SHAPES_SQUEEZE),
    'resnet50': (1242, 375, 20, 24, 78, _SHAPES_RES50),
}


def _kitti_config(net):
  width, height, batch, gh, gw, shapes = _KITTI_NETS[net]
  mc = base_model_config('KITTI')
  mc.update(
      IMAGE_WIDTH=width, IMAGE_HEIGHT=height, BATCH_SIZE=batch,
      WEIGHT_DECAY=0.0001, LEARNING_RATE=0.01, DECAY_STEPS=10000,
      MAX_GRAD_NORM=1.0, MOMENTUM=0.9, LR_DECAY_FACTOR=0.5,
      LOSS_COEF_BBOX=5.0, LOSS_COEF_CONF_POS=75.0, LOSS_COEF_CONF_NEG=100.0,
      LOSS_COEF_CLASS=1.0,
      PLOT_PROB_THRESH=0.4, NMS_THRESH=0.4, PROB_THRESH=0.005, TOP_N_DETECTION=64,
      DATA_AUGMENTATION=True, DRIFT_X=150, DRIFT_Y=100,
      EXCLUDE_HARD_EXAMPLES=False)
  mc.GRID_H, mc.GRID_W = gh, gw          # engine extras (not in the reference mc)
  mc.ANCHOR_SHAPES = shapes
  mc.ANCHOR_BOX = make_anchor_box(width, height, gh, gw, shapes)
  mc.ANCHORS = len(mc.ANCHOR_BOX)
  mc.ANCHOR_PER_GRID = len(shapes)
  return mc


def set_anchors(mc):
  """Recompute ``mc.ANCHOR_BOX`` after IMAGE_WIDTH/HEIGHT were overridden (the
  reference's per-config ``set_anchors(mc)``)."""
  return make_anchor_box(mc.IMAGE_WIDTH, mc.IMAGE_HEIGHT, mc.GRID_H, mc.GRID_W,
                         mc.ANCHOR_SHAPES)


def kitti_squeezeDet_config():
  return _kitti_config('squeezeDet')


def kitti_squeezeDetPlus_config():
  return _kitti_config('squeezeDet+')


def kitti_vgg16_config():
  return _kitti_config('vgg16')


def kitti_res50_config():
  return _kitti_config('resnet50')
