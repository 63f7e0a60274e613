"""TF-1.0 NHWC op semantics restated in numpy (oracle — test infrastructure only).

Follows (reference, read-only):
  * ``src/nn_skeleton.py:471-563``  ``_conv_layer``   -> conv2d + bias + relu
  * ``src/nn_skeleton.py:565-586``  ``_pooling_layer`` -> tf.nn.max_pool
  * ``src/nn_skeleton.py:374-468``  ``_conv_bn_layer`` -> conv2d [+bias] + frozen BN
The arithmetic itself lives in tensorflow-gpu==1.0.0 (requirements.txt:6, not
vendored): the geometry below is TF's documented SAME/VALID rule
(SURVEY.md App. A.1); parity of this half is UNPINNED (no TF, no golden
vectors) and is cross-checked naive-loops vs im2col-GEMM vs torch-CPU in tests.
"""
from __future__ import annotations

import numpy as np


def conv_geometry(in_size: int, k: int, stride: int, padding: str):
  """Output size and (pad_before, pad_after) for one spatial dim.

  TF rule: SAME  -> out = ceil(in/stride), pad_total = max((out-1)*s + k - in, 0),
                    pad_before = pad_total // 2 (the extra cell goes after).
           VALID -> out = floor((in-k)/s) + 1, no padding.
  """
  padding = padding.upper()
  if padding == 'SAME':
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2, total - total // 2
  if padding == 'VALID':
    return (in_size - k) // stride + 1, 0, 0
  raise ValueError('padding must be SAME or VALID, got %r' % (padding,))


def relu(x):
  return np.maximum(x, 0)


def conv2d(x, w, b=None, stride=1, padding='SAME', apply_relu=False,
           dtype=np.float32, rows_per_chunk=64):
  """``relu?(conv2d(x, w, [1,s,s,1], padding) + b)``  (nn_skeleton.py:539-547).

  x [B,H,W,Cin] NHWC; w [kh,kw,Cin,Cout] (HWIO, nn_skeleton.py:531-533);
  cross-correlation (no kernel flip); zero padding.  im2col + GEMM in `dtype`.
  """
  x = np.asarray(x, dtype=dtype)
  w = np.asarray(w, dtype=dtype)
  B, H, W, C = x.shape
  kh, kw, ci, co = w.shape
  assert ci == C, (ci, C)
  Ho, pt, pb = conv_geometry(H, kh, stride, padding)
  Wo, pl, pr = conv_geometry(W, kw, stride, padding)
  if pt or pb or pl or pr:
    x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
  wmat = w.reshape(kh * kw * C, co)
  out = np.empty((B, Ho, Wo, co), dtype=dtype)
  for n in range(B):
    if kh == 1 and kw == 1:
      xs = x[n, ::stride, ::stride][:Ho, :Wo]
      out[n] = (xs.reshape(-1, C) @ wmat).reshape(Ho, Wo, co)
      continue
    win = np.lib.stride_tricks.sliding_window_view(x[n], (kh, kw), axis=(0, 1))
    # win [Hp-kh+1, Wp-kw+1, C, kh, kw] -> strided -> [Ho, Wo, kh, kw, C]
    win = win[::stride, ::stride][:Ho, :Wo].transpose(0, 1, 3, 4, 2)
    for r0 in range(0, Ho, rows_per_chunk):
      r1 = min(Ho, r0 + rows_per_chunk)
      cols = np.ascontiguousarray(win[r0:r1]).reshape(-1, kh * kw * C)
      out[n, r0:r1] = (cols @ wmat).reshape(r1 - r0, Wo, co)
  if b is not None:
    out += np.asarray(b, dtype=dtype)
  if apply_relu:
    np.maximum(out, 0, out=out)
  return out


def conv2d_naive(x, w, b=None, stride=1, padding='SAME', apply_relu=False,
                 dtype=np.float64):
  """Direct-loop conv for tiny shapes: the independent check on ``conv2d``."""
  x = np.asarray(x, dtype=dtype)
  w = np.asarray(w, dtype=dtype)
  B, H, W, C = x.shape
  kh, kw, _, co = w.shape
  Ho, pt, _ = conv_geometry(H, kh, stride, padding)
  Wo, pl, _ = conv_geometry(W, kw, stride, padding)
  out = np.zeros((B, Ho, Wo, co), dtype=dtype)
  for n in range(B):
    for i in range(Ho):
      for j in range(Wo):
        acc = np.zeros(co, dtype=dtype)
        for u in range(kh):
          y = i * stride + u - pt
          if y < 0 or y >= H:
            continue
          for v in range(kw):
            xx = j * stride + v - pl
            if xx < 0 or xx >= W:
              continue
            acc += x[n, y, xx] @ w[u, v]
        out[n, i, j] = acc
  if b is not None:
    out += np.asarray(b, dtype=dtype)
  if apply_relu:
    out = np.maximum(out, 0)
  return out


def max_pool(x, k, stride, padding='SAME'):
  """``tf.nn.max_pool`` NHWC (nn_skeleton.py:580-583).  SAME never reads the
  padding: padded cells are excluded from the max (equivalently -inf)."""
  x = np.asarray(x)
  B, H, W, C = x.shape
  Ho, pt, pb = conv_geometry(H, k, stride, padding)
  Wo, pl, pr = conv_geometry(W, k, stride, padding)
  if pt or pb or pl or pr:
    x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)),
               constant_values=-np.inf)
  win = np.lib.stride_tricks.sliding_window_view(x, (k, k), axis=(1, 2))
  win = win[:, ::stride, ::stride][:, :Ho, :Wo]
  return win.max(axis=(-2, -1))


def max_pool_naive(x, k, stride, padding='SAME'):
  x = np.asarray(x)
  B, H, W, C = x.shape
  Ho, pt, _ = conv_geometry(H, k, stride, padding)
  Wo, pl, _ = conv_geometry(W, k, stride, padding)
  out = np.full((B, Ho, Wo, C), -np.inf, dtype=x.dtype)
  for i in range(Ho):
    for j in range(Wo):
      for u in range(k):
        y = i * stride + u - pt
        if not 0 <= y < H:
          continue
        for v in range(k):
          xx = j * stride + v - pl
          if 0 <= xx < W:
            out[:, i, j] = np.maximum(out[:, i, j], x[:, y, xx])
  return out


def batch_norm_frozen(x, mean, var, beta, gamma, eps=1e-5):
  """``tf.nn.batch_normalization`` with stored statistics
  (nn_skeleton.py:447-449; eps = mc.BATCH_NORM_EPSILON, config.py:131).
  TF 1.0 evaluates it as  inv = rsqrt(var+eps)*gamma ; x*inv + (beta - mean*inv)."""
  dt = x.dtype
  inv = (1.0 / np.sqrt(np.asarray(var, dt) + dt.type(eps))) * np.asarray(gamma, dt)
  return x * inv + (np.asarray(beta, dt) - np.asarray(mean, dt) * inv)
