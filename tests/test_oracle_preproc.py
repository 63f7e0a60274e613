"""Pins oracle/preproc.py against the installed cv2 (the reference's own dependency for this
step): float32 INTER_LINEAR resize in both orderings of the mean subtraction."""
import numpy as np
import pytest

from oracle import preproc

cv2 = pytest.importorskip('cv2')

MEANS = np.array([[[103.939, 116.779, 123.68]]])
ULP255 = float(np.spacing(np.float32(255.0)))

SHAPES = [(370, 1224, 375, 1242),    # a KITTI frame size -> mc.IMAGE_*
          (375, 1242, 375, 1242),    # identity
          (720, 1280, 375, 1242), (100, 150, 375, 1242), (480, 640, 384, 1248), (37, 41, 19, 23),
          (5, 7, 31, 3)]


@pytest.mark.parametrize('h0,w0,h,w', SHAPES)
def test_resize_restatement_within_3ulp_of_cv2(h0, w0, h, w):
  rng = np.random.default_rng(h0 * 7 + w)
  im = rng.integers(0, 256, (h0, w0, 3), dtype=np.uint8)
  want = cv2.resize(im.astype(np.float32), (w, h))
  got = preproc.resize_linear_f32(im, w, h)
  assert got.shape == want.shape and got.dtype == np.float32
  assert np.abs(got - want).max() <= 3 * ULP255
  if (h0, w0) == (h, w):
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('order', ['demo', 'eval'])
def test_preprocess_orderings_match_the_reference_expressions(order):
  rng = np.random.default_rng(3)
  im = rng.integers(0, 256, (370, 1224, 3), dtype=np.uint8)
  if order == 'demo':                                   # src/demo.py:187-190
    x = im.astype(np.float32, copy=False)
    x = cv2.resize(x, (1242, 375))
    want = (x - MEANS).astype(np.float32)
  else:                                                 # src/dataset/imdb.py:87-91
    x = im.astype(np.float32, copy=False)
    x -= MEANS
    want = cv2.resize(x, (1242, 375))
  got = preproc.preprocess(im, 1242, 375, MEANS, order)
  assert np.abs(got - want).max() <= 3 * ULP255
