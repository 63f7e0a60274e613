"""CPU restatement of the image pre-processing in front of the hot path — test
infrastructure, like the rest of oracle/.

Reference: demo (src/demo.py:187-190) ``im = im.astype(np.float32); im = cv2.resize(im, (W, H));
input = im - mc.BGR_MEANS``; evaluation (src/dataset/imdb.py:85-97) ``im = im.astype(np.float32);
im -= mc.BGR_MEANS; im = cv2.resize(im, (W, H))``.

`cv2.resize` on float32 with the default INTER_LINEAR is restated from OpenCV's resize.cpp
(cv::resize -> resizeGeneric_, HResizeLinear / VResizeLinear), PINNED here against the installed
cv2 (4.13): sampling position ``f = (d + 0.5) * (1 / (dst / src)) - 0.5`` in double, tap
``s = floor(f)``, weight ``float32(f - s)``; columns clamp with weight 0 (``s < 0`` or
``s >= src - 1``), rows clamp the two row indices and keep their weights; horizontal pass then
vertical pass in float32, ``x0 * (1 - w) + x1 * w``.  OpenCV's SIMD build contracts some of these
into FMAs, so the restatement is within 3 float32 ulp of cv2 (tests/test_oracle_preproc.py), not
bit-equal.  (OpenCV <= 3.x, the reference's era, rounded ``f`` to float32 before taking the floor;
that variant differs from this one by up to 6e-5 of the pixel range.)"""
import numpy as np


def linear_coeffs(dst, src):
  inv = np.float64(dst) / np.float64(src)
  scale = 1.0 / inv
  f = (np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5
  s = np.floor(f).astype(np.int64)
  return s, (f - s).astype(np.float32)


def resize_linear_f32(img, width, height):
  """cv2.resize(img.astype(float32), (width, height)) for an [H0, W0, C] image."""
  img = np.asarray(img, np.float32)
  h0, w0 = img.shape[:2]
  if (h0, w0) == (height, width):
    return img.copy()
  sx, fx = linear_coeffs(width, w0)
  sy, fy = linear_coeffs(height, h0)
  lo = sx < 0
  fx = np.where(lo, np.float32(0), fx)
  sx = np.where(lo, 0, sx)
  hi = sx >= w0 - 1
  fx = np.where(hi, np.float32(0), fx)
  sx = np.where(hi, w0 - 1, sx)
  sx1 = np.minimum(sx + 1, w0 - 1)
  a0 = (np.float32(1) - fx)[None, :, None]
  a1 = fx[None, :, None]
  rows = img[:, sx, :] * a0 + img[:, sx1, :] * a1
  rows = np.where(hi[None, :, None], img[:, sx, :], rows)
  y0 = np.clip(sy, 0, h0 - 1)
  y1 = np.clip(sy + 1, 0, h0 - 1)
  b0 = (np.float32(1) - fy)[:, None, None]
  b1 = fy[:, None, None]
  return (rows[y0] * b0 + rows[y1] * b1).astype(np.float32)


def preprocess(img_u8, width, height, bgr_means, order='demo'):
  """uint8 BGR image -> the float32 [height, width, 3] tensor the reference feeds.
  order 'demo': resize, then subtract the means (float64 subtraction, float32 at the feed);
  order 'eval': subtract in place in float32, then resize."""
  means = np.asarray(bgr_means, np.float64).reshape(1, 1, 3)
  im = np.asarray(img_u8).astype(np.float32)
  if order == 'demo':
    return (resize_linear_f32(im, width, height) - means).astype(np.float32)
  if order == 'eval':
    im = (im - means).astype(np.float32)              # `im -= BGR_MEANS` on a float32 array
    return resize_linear_f32(im, width, height)
  raise ValueError(order)
