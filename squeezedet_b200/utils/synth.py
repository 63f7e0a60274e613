"""Synthetic, seeded inputs for benchmarking and parity runs (no dataset or
checkpoint is reachable offline).  The reference's own initialisers
(truncated normal, stddev 1e-3/1e-2/1e-4, zero bias — nn_skeleton.py:527-529)
produce |preds| ~ 1e-23 and 16 848 identical scores, useless for parity; these
keep every pre-activation O(1) analytically (He scaling, first layer divided by
the input's standard deviation) so scores, classes and boxes are well spread."""
from __future__ import annotations

import json
import os

import numpy as np

BGR_MEANS = np.array([[[103.939, 116.779, 123.68]]])
INPUT_STD = 74.0          # std of U{0..255} minus its mean


def synthetic_images(batch, height, width, seed=1234):
  """uint8 ~ U{0..255} -> float32 -> minus BGR means (the demo.py:187-190 recipe
  without a file), shape [batch, height, width, 3]."""
  rng = np.random.default_rng(seed)
  img = rng.integers(0, 256, size=(batch, height, width, 3), dtype=np.uint8)
  return (img.astype(np.float32) - BGR_MEANS.astype(np.float32)).astype(np.float32)


_GAINS_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'synth_gains.json')
_gains_cache = None


def layer_gains(net=None, specs=None):
  """Committed per-layer gains (made by tests/golden/make_calibration.py) that bring
  every pre-activation to unit std (ConvDet head: 1.5).  The net is recognised from
  the parameter list when not named."""
  global _gains_cache
  if _gains_cache is None:
    try:
      with open(_GAINS_FILE) as f:
        _gains_cache = json.load(f)
    except OSError:
      _gains_cache = {}
  if net is None and specs is not None:
    scopes = {n.rsplit('/', 1)[0] for n, _ in specs}
    for cand, table in _gains_cache.items():
      if scopes == set(table):
        # squeezeDet and squeezeDet+ share scope names: tell them apart by conv1's size
        k = dict(specs).get('conv1/kernels')
        if cand == 'squeezeDet' and k is not None and k[0] != 3:
          continue
        if cand == 'squeezeDet+' and k is not None and k[0] != 7:
          continue
        return table
    return {}
  return _gains_cache.get(net, {})


def synthetic_weights(param_specs, seed=0, head_std=1.5, head_names=None, gains=None):
  """param_specs: iterable of (name, shape) in reference naming.  Returns
  {name: float32 ndarray}.  `head_names`: scopes of the ConvDet head (default: the
  last '<x>/kernels' in the list).  `gains`: {scope: multiplier} applied to the
  scope's kernels+biases (BN scopes: gamma+beta); default = the committed table."""
  specs = [(n, tuple(int(s) for s in shp)) for n, shp in param_specs]
  kernels = [n for n, s in specs if n.endswith('/kernels')]
  if head_names is None:
    head_names = {kernels[-1].rsplit('/', 1)[0]}
  first = kernels[0].rsplit('/', 1)[0]
  if gains is None:
    gains = layer_gains(specs=specs)
  bn_scopes = {n.rsplit('/', 1)[0] for n, _ in specs if n.endswith('/gamma')}
  rng = np.random.default_rng(seed)
  out = {}
  for name, shape in specs:
    scope, leaf = name.rsplit('/', 1)
    g = gains.get(scope, 1.0)
    scaled = ('gamma', 'beta') if scope in bn_scopes else ('kernels', 'biases')
    if leaf == 'kernels':
      fan_in = shape[0] * shape[1] * shape[2]
      if scope == first:
        std = np.sqrt(1.0 / fan_in) / INPUT_STD
      elif scope in head_names:
        std = head_std * np.sqrt(2.0 / fan_in)
      else:
        std = np.sqrt(2.0 / fan_in)
      val = rng.normal(0.0, std, size=shape)
    elif leaf == 'biases':
      val = rng.normal(0.0, 0.1, size=shape)
    elif leaf == 'gamma':
      val = rng.uniform(0.5, 1.5, size=shape)
    elif leaf == 'beta':
      val = rng.normal(0.0, 0.1, size=shape)
    elif leaf == 'mean':
      val = rng.normal(0.0, 0.1, size=shape)
    elif leaf == 'var':
      val = rng.uniform(0.5, 1.5, size=shape)
    else:
      raise ValueError('unknown parameter kind: ' + name)
    if leaf in scaled:
      val = val * g
    out[name] = val.astype(np.float32)
  return out


def model_param_specs(model):
  return [(p.name, p.shape) for p in model.model_params]
