"""Weight files for the engine.  No TensorFlow checkpoint reader exists yet (SURVEY §8 f-2,
round 2): weights are exchanged as ``.npz`` archives keyed by the reference's TF variable names
(`conv1/kernels` HWIO, `fire2/squeeze1x1/biases`, BN `.../gamma|beta|mean|var`) — exactly the
names `tf.train.Saver(model.model_params)` stores (reference src/demo.py:181) — or as the
reference's Caffe-derived joblib ``.pkl`` ({layer: [W(out,in,h,w), b]}, src/nn_skeleton.py:492-508)."""
from __future__ import annotations

import numpy as np


def load_npz(path):
  with np.load(path) as z:
    return {k: z[k] for k in z.files}


def save_npz(path, weights):
  np.savez(path, **{k: np.asarray(v, dtype=np.float32) for k, v in weights.items()})


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
