"""The TF half of the oracle (parity UNPINNED: no TF, no reference vectors) is
cross-validated three ways: naive fp64 loops vs im2col-GEMM vs torch-CPU, plus the
shape/FLOP/parameter tables of SURVEY App. B (reference counter formulas,
src/nn_skeleton.py:549-561)."""
import numpy as np
import pytest

import oracle
from oracle import semantics as S


def test_geometry_matches_survey_appendix():
  # @375x1242: conv1 3x3/2 SAME -> 188x621 pads H(1,1) W(0,1)
  assert S.conv_geometry(375, 3, 2, 'SAME') == (188, 1, 1)
  assert S.conv_geometry(1242, 3, 2, 'SAME') == (621, 0, 1)
  assert S.conv_geometry(188, 3, 2, 'SAME') == (94, 0, 1)      # pool1
  assert S.conv_geometry(621, 3, 2, 'SAME') == (311, 1, 1)
  assert S.conv_geometry(47, 3, 2, 'SAME') == (24, 1, 1)       # pool5
  assert S.conv_geometry(156, 3, 2, 'SAME') == (78, 0, 1)
  assert S.conv_geometry(384, 3, 2, 'SAME') == (192, 0, 1)     # @384x1248 always (0,1)
  assert S.conv_geometry(375, 7, 2, 'VALID') == (185, 0, 0)    # squeezeDet+ conv1
  assert S.conv_geometry(1242, 7, 2, 'VALID') == (618, 0, 0)


@pytest.mark.parametrize('k,stride,padding,H,W,Cin,Cout', [
    (3, 2, 'SAME', 9, 12, 3, 5), (3, 1, 'SAME', 7, 8, 4, 6), (1, 1, 'SAME', 5, 6, 8, 3),
    (7, 2, 'VALID', 15, 17, 3, 4), (1, 2, 'SAME', 9, 10, 4, 4), (3, 2, 'SAME', 8, 11, 2, 3)])
def test_conv_im2col_vs_naive_vs_torch(k, stride, padding, H, W, Cin, Cout):
  import torch
  import torch.nn.functional as F
  rng = np.random.default_rng(k * 100 + stride)
  x = rng.normal(size=(2, H, W, Cin))
  w = rng.normal(size=(k, k, Cin, Cout))
  b = rng.normal(size=(Cout,))
  ref = S.conv2d_naive(x, w, b, stride, padding, apply_relu=True)
  got64 = S.conv2d(x, w, b, stride, padding, apply_relu=True, dtype=np.float64)
  np.testing.assert_allclose(got64, ref, rtol=1e-12, atol=1e-12)
  got32 = S.conv2d(x, w, b, stride, padding, apply_relu=True, dtype=np.float32)
  np.testing.assert_allclose(got32, ref, rtol=2e-5, atol=2e-5)
  # torch: explicit asymmetric TF padding then a VALID conv
  _, pt, pb = S.conv_geometry(H, k, stride, padding)
  _, pl, pr = S.conv_geometry(W, k, stride, padding)
  xt = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (pl, pr, pt, pb))
  yt = F.conv2d(xt, torch.from_numpy(w).permute(3, 2, 0, 1), torch.from_numpy(b), stride)
  yt = torch.relu(yt).permute(0, 2, 3, 1).numpy()
  np.testing.assert_allclose(yt, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('k,stride,padding,H,W', [
    (3, 2, 'SAME', 9, 12), (3, 2, 'VALID', 9, 12), (2, 2, 'SAME', 7, 9), (3, 2, 'SAME', 8, 11)])
def test_max_pool_vs_naive_vs_torch(k, stride, padding, H, W):
  import torch
  import torch.nn.functional as F
  rng = np.random.default_rng(3)
  x = rng.normal(size=(2, H, W, 5)).astype(np.float32) - 3.0     # all-negative regions too
  ref = S.max_pool_naive(x, k, stride, padding)
  got = S.max_pool(x, k, stride, padding)
  assert np.array_equal(got, ref)
  _, pt, pb = S.conv_geometry(H, k, stride, padding)
  _, pl, pr = S.conv_geometry(W, k, stride, padding)
  xt = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float('-inf'))
  yt = F.max_pool2d(xt, k, stride).permute(0, 2, 3, 1).numpy()
  assert np.array_equal(yt, ref)


def test_batch_norm_frozen():
  rng = np.random.default_rng(0)
  x = rng.normal(size=(2, 3, 4, 5))
  mean, var = rng.normal(size=5), rng.uniform(0.5, 1.5, 5)
  beta, gamma = rng.normal(size=5), rng.uniform(0.5, 1.5, 5)
  want = (x - mean) * gamma / np.sqrt(var + 1e-5) + beta
  np.testing.assert_allclose(S.batch_norm_frozen(x, mean, var, beta, gamma, 1e-5), want,
                             rtol=1e-12)


@pytest.mark.parametrize('net,hw,gflop,params,grid', [
    ('squeezeDet', (375, 1242), 10.4923, 2082120, (24, 78)),
    ('squeezeDet', (384, 1248), 10.5661, 2082120, (24, 78)),
    ('squeezeDet+', (375, 1242), 77.068, 7021640, (22, 76)),
    ('resnet50', (375, 1242), 61.127, None, (24, 78)),
    ('vgg16', (375, 1242), 288.042, None, (24, 78))])
def test_layer_tables_match_survey_appendix_b(net, hw, gflop, params, grid):
  rows = oracle.layer_table(net, hw[0], hw[1])
  assert abs(sum(r[3] for r in rows) / 1e9 - gflop) < 2e-3
  if params is not None:
    assert sum(r[4] for r in rows) == params
  assert rows[-1][2] == (grid[0], grid[1], 72)


def test_forward_tiny_squeezedet_runs_and_fp32_tracks_fp64():
  from squeezedet_b200.utils import synth
  specs = oracle.param_specs('squeezeDet')
  w = synth.synthetic_weights(specs, seed=1)
  x = synth.synthetic_images(1, 48, 80, seed=2)
  p32 = oracle.forward('squeezeDet', w, x, dtype=np.float32)
  p64 = oracle.forward('squeezeDet', w, x, dtype=np.float64)
  assert p32.shape == (1, 3, 5, 72)
  assert np.isfinite(p64).all() and 0.2 < p64.std() < 20
  np.testing.assert_allclose(p32, p64, rtol=0, atol=2e-4 * np.abs(p64).max())
