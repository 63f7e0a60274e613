"""Weight files for the engine, all keyed by the reference's TF variable names (`conv1/kernels`
HWIO, `fire2/squeeze1x1/biases`, BN `.../gamma|beta|mean|var`) — exactly the names
`tf.train.Saver(model.model_params)` stores (reference src/demo.py:181):

* TensorFlow checkpoints, V2 bundles and V1 tables, read without TensorFlow
  (`tf_checkpoint.py`; also written, so the reference's Saver can restore our weights);
* ``.npz`` archives;
* the reference's Caffe-derived joblib ``.pkl`` ({layer: [W(out,in,h,w), b]},
  src/nn_skeleton.py:492-508)."""
from __future__ import annotations

import os

import numpy as np


def load_npz(path):
  with np.load(path) as z:
    return {k: z[k] for k in z.files}


def save_npz(path, weights):
  np.savez(path, **{k: np.asarray(v, dtype=np.float32) for k, v in weights.items()})


def load_weights_file(path, names=None):
  """{variable name: ndarray} from whatever `--checkpoint` points at: an .npz, or the Saver path
  of a TensorFlow checkpoint (``.../model.ckpt-87000``, as in reference src/demo.py:181-184).
  `names` (e.g. ``model.param_names()``) restricts a TensorFlow checkpoint to the variables the
  model restores, like ``Saver(model.model_params)`` does - the optimizer slots and counters
  stored next to them are then neither read nor checksummed."""
  from . import tf_checkpoint
  if path.endswith('.npz'):
    return load_npz(path)
  if tf_checkpoint.checkpoint_kind(path):
    return tf_checkpoint.read_checkpoint(path, names=set(names) if names is not None else None)
  if os.path.exists(path + '.npz'):
    return load_npz(path + '.npz')
  raise FileNotFoundError('%s: no .npz archive and no TensorFlow checkpoint (V2 .index/.data or '
                          'V1) at this path' % path)


def save_tf_checkpoint(prefix, weights):
  """Write a V2 checkpoint the reference's `saver.restore(sess, prefix)` accepts."""
  from . import tf_checkpoint
  tf_checkpoint.write_v2(prefix, {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()})


# tf defaults of a _conv_bn_layer built without pretrained blobs (nn_skeleton.py:412-420)
BN_DEFAULTS = {'gamma': 1.0, 'beta': 0.0, 'mean': 0.0, 'var': 1.0}


def caffe_bn_names(conv_name):
  """Caffe blob names of a ResNet conv's BatchNorm / Scale layers, as the reference passes them
  to _conv_bn_layer (src/nets/resnet50_convDet.py:41-43, :140-168): 'conv1' -> ('bn_conv1',
  'scale_conv1'); 'res2a_branch2a' -> ('bn2a_branch2a', 'scale2a_branch2a')."""
  if conv_name.startswith('res'):
    return 'bn' + conv_name[3:], 'scale' + conv_name[3:]
  return 'bn_' + conv_name, 'scale_' + conv_name


def from_caffe_pkl(path, model, blobs=None):
  """ImageNet-pretrained Caffe blobs (joblib pkl) -> {reference variable name: ndarray}.
  Kernels are stored [out, in, h, w] and transposed to HWIO like the reference does
  (nn_skeleton.py:496, :405); BatchNorm statistics come from the `bn*` blob (mean, var) and
  gamma / beta from the `scale*` blob (nn_skeleton.py:408-411).  Layers whose shapes do not match
  are skipped with a message, as the reference prints (:499-508); BN parameters that have no
  blob get the reference's initialiser values (gamma 1, beta 0, mean 0, var 1; :416-420) so a
  partial import can never leave a BN layer multiplying by zero."""
  if blobs is None:
    import joblib
    blobs = joblib.load(path)
  out = {}
  for p in model.model_params:
    scope, leaf = p.name.rsplit('/', 1)
    layer = scope.split('/')[-1] if scope not in blobs else scope
    val = None
    if leaf in ('kernels', 'biases'):
      if layer in blobs:
        val = np.asarray(blobs[layer][0 if leaf == 'kernels' else 1])
        if leaf == 'kernels':
          val = np.transpose(val, [2, 3, 1, 0])
    elif leaf in BN_DEFAULTS:
      bn_name, scale_name = caffe_bn_names(layer)
      src, idx = {'mean': (bn_name, 0), 'var': (bn_name, 1),
                  'gamma': (scale_name, 0), 'beta': (scale_name, 1)}[leaf]
      if src in blobs:
        val = np.asarray(blobs[src][idx])
      else:
        val = np.full(p.shape, BN_DEFAULTS[leaf], np.float32)
    if val is None:
      continue
    if leaf != 'kernels':
      val = np.asarray(val).reshape(-1)        # Caffe stores per-channel blobs as 1-D or [1,C,1,1]
    if tuple(val.shape) != tuple(p.shape):
      print('Shape of the pretrained parameter of {} does not match, skipped'.format(p.name))
      if leaf in BN_DEFAULTS:
        out[p.name] = np.full(p.shape, BN_DEFAULTS[leaf], np.float32)
      continue
    out[p.name] = val.astype(np.float32)
  return out
