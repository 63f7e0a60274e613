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


def from_caffe_pkl(path, model):
  """ImageNet-pretrained Caffe blobs (joblib pkl): kernels are stored [out, in, h, w] and
  transposed to HWIO like the reference does (nn_skeleton.py:496); layers whose shapes do not
  match are skipped with a message, as the reference prints (:499-508)."""
  import joblib
  blobs = joblib.load(path)
  out = {}
  for p in model.model_params:
    scope, leaf = p.name.rsplit('/', 1)
    layer = scope.split('/')[-1] if scope not in blobs else scope
    if layer not in blobs or leaf not in ('kernels', 'biases'):
      continue
    val = np.asarray(blobs[layer][0 if leaf == 'kernels' else 1])
    if leaf == 'kernels':
      val = np.transpose(val, [2, 3, 1, 0])
    if tuple(val.shape) != tuple(p.shape):
      print('Shape of the pretrained parameter of {} does not match, skipped'.format(p.name))
      continue
    out[p.name] = val.astype(np.float32)
  return out
