"""Host-side pieces of the demo/eval entry points (no GPU): flags, preprocessing order, KITTI
line format, box drawing."""
import numpy as np

from squeezedet_b200 import demo, eval as sq_eval
from squeezedet_b200 import config as cfg
from squeezedet_b200.utils import viz
from squeezedet_b200._lib import DET_DTYPE


def test_demo_flags_match_reference_names():
  f = demo.parse_flags(['--mode', 'image', '--input_path', 'x/*.png', '--out_dir', 'o',
                        '--demo_net', 'squeezeDet+', '--gpu', '1', '--checkpoint', 'synthetic'])
  assert (f.mode, f.input_path, f.out_dir, f.demo_net, f.gpu, f.checkpoint) == \
      ('image', 'x/*.png', 'o', 'squeezeDet+', '1', 'synthetic')
  e = sq_eval.parse_flags(['--net', 'resnet50', '--image_set', 'val', '--data_path', '/d'])
  assert (e.net, e.image_set, e.data_path, e.dataset) == ('resnet50', 'val', '/d', 'KITTI')


def test_demo_preprocess_resizes_then_subtracts_mean():
  mc = cfg.kitti_squeezeDet_config()
  rng = np.random.default_rng(0)
  im = rng.integers(0, 256, (375, 1242, 3), dtype=np.uint8)
  shown, fed = demo.preprocess(im, mc)
  assert shown.shape == (384, 1248, 3) and fed.shape == (384, 1248, 3) and fed.dtype == np.float32
  np.testing.assert_allclose(fed, (shown - mc.BGR_MEANS).astype(np.float32))


def test_kitti_line_format():
  line = viz.kitti_detection_line('Car', [1.0, 2.345, 300.999, 40.0], 0.98765)
  assert line == 'car -1 -1 0.0 1.00 2.35 301.00 40.00 0.0 0.0 0.0 0.0 0.0 0.0 0.0 0.988\n' or \
      line == 'car -1 -1 0.0 1.00 2.34 301.00 40.00 0.0 0.0 0.0 0.0 0.0 0.0 0.0 0.988\n'
  assert len(line.split()) == 16


def test_write_and_parse_kitti_files(tmp_path):
  all_boxes = [[[[1, 2, 3, 4, 0.5]], []], [[], [[5, 6, 7, 8, 0.25], [9, 10, 11, 12, 0.125]]]]
  d = viz.write_kitti_detections(str(tmp_path / 'det' / 'data'), ['000001', '000002'],
                                 ('car', 'pedestrian'), all_boxes)
  assert open(tmp_path / 'det' / 'data' / '000001.txt').read().startswith('car -1 -1 0.0 1.00')
  assert len(open(tmp_path / 'det' / 'data' / '000002.txt').read().splitlines()) == 2
  (tmp_path / 'det' / 'stats_car_ap.txt').write_text('a = 0.9\nb = 0.8\nc = 0.7\n')
  aps, names = viz.parse_kitti_ap_files(d, ('car', 'pedestrian'))
  assert aps == [0.9, 0.8, 0.7, 0.0, 0.0, 0.0] and names[0] == 'car_easy' and names[-1] == 'pedestrian_hard'


def test_detections_to_all_boxes_rescales_and_converts():
  recs = np.zeros(3, DET_DTYPE)
  recs[0] = (5, 0, 0.9, 100.0, 50.0, 20.0, 10.0)
  recs[1] = (9, 2, 0.4, 10.0, 10.0, 4.0, 4.0)
  out = sq_eval.detections_to_all_boxes(recs, 2, (2.0, 0.5), 3)
  assert len(out[0]) == 1 and len(out[1]) == 0 and len(out[2]) == 1
  np.testing.assert_allclose(out[0][0], [45.0, 90.0, 55.0, 110.0, 0.9], rtol=1e-6)


def test_draw_box_marks_pixels():
  im = np.zeros((60, 80, 3), np.float32)
  viz.draw_box(im, [np.array([40., 30., 20., 10.], np.float32)], ['car: (0.90)'],
               cdict=viz.CLASS_COLORS)
  assert tuple(im[25, 30]) == (255.0, 191.0, 0.0)          # top-left corner, class colour
  assert im.sum() > 0


def test_caffe_pkl_import_fills_batchnorm_like_the_reference():
  """reference nn_skeleton.py:403-411: kernels from cw[conv], mean/var from cw[bn_*],
  gamma/beta from cw[scale_*]; missing BN blobs fall back to the tf initialisers (1,0,0,1),
  never to zeros."""
  from squeezedet_b200.utils import checkpoint as ckpt

  class P:
    def __init__(self, name, shape):
      self.name, self.shape = name, shape

  class M:
    model_params = [
        P('conv1/kernels', (7, 7, 3, 4)), P('conv1/biases', (4,)),
        P('conv1/gamma', (4,)), P('conv1/beta', (4,)), P('conv1/mean', (4,)), P('conv1/var', (4,)),
        P('conv2_x/res2a/res2a_branch1/kernels', (1, 1, 4, 8)),
        P('conv2_x/res2a/res2a_branch1/gamma', (8,)), P('conv2_x/res2a/res2a_branch1/beta', (8,)),
        P('conv2_x/res2a/res2a_branch1/mean', (8,)), P('conv2_x/res2a/res2a_branch1/var', (8,)),
        P('conv2_x/res2b/res2b_branch2/res2b_branch2a/kernels', (1, 1, 8, 2)),
        P('conv2_x/res2b/res2b_branch2/res2b_branch2a/gamma', (2,)),
        P('conv2_x/res2b/res2b_branch2/res2b_branch2a/var', (2,)),
    ]

  rng = np.random.default_rng(0)
  blobs = {
      'conv1': [rng.normal(size=(4, 3, 7, 7)), rng.normal(size=(4,))],
      'bn_conv1': [rng.normal(size=(4,)), rng.uniform(0.5, 2, size=(4,))],
      'scale_conv1': [rng.normal(size=(4,)), rng.normal(size=(4,))],
      'res2a_branch1': [rng.normal(size=(8, 4, 1, 1))],
      'bn2a_branch1': [rng.normal(size=(1, 8, 1, 1)), rng.uniform(0.5, 2, size=(8,))],
      'scale2a_branch1': [rng.normal(size=(8,)), rng.normal(size=(8,))],
      'res2b_branch2a': [rng.normal(size=(2, 8, 1, 1))],
      # no bn2b_branch2a / scale2b_branch2a blobs
  }
  w = ckpt.from_caffe_pkl(None, M, blobs=blobs)
  np.testing.assert_allclose(w['conv1/kernels'], np.transpose(blobs['conv1'][0], [2, 3, 1, 0]),
                             rtol=1e-6)
  np.testing.assert_allclose(w['conv1/mean'], blobs['bn_conv1'][0], rtol=1e-6)
  np.testing.assert_allclose(w['conv1/var'], blobs['bn_conv1'][1], rtol=1e-6)
  np.testing.assert_allclose(w['conv1/gamma'], blobs['scale_conv1'][0], rtol=1e-6)
  np.testing.assert_allclose(w['conv1/beta'], blobs['scale_conv1'][1], rtol=1e-6)
  np.testing.assert_allclose(w['conv2_x/res2a/res2a_branch1/mean'],
                             blobs['bn2a_branch1'][0].reshape(-1), rtol=1e-6)
  assert ckpt.caffe_bn_names('res4f_branch2c') == ('bn4f_branch2c', 'scale4f_branch2c')
  g = w['conv2_x/res2b/res2b_branch2/res2b_branch2a/gamma']
  v = w['conv2_x/res2b/res2b_branch2/res2b_branch2a/var']
  assert np.all(g == 1.0) and np.all(v == 1.0)


def test_all_boxes_rejects_overflow_marker():
  import pytest
  from squeezedet_b200._lib import SqdetError
  recs = np.zeros(3, DET_DTYPE)
  with pytest.raises(SqdetError):
    sq_eval.detections_to_all_boxes(recs, -1, None, 3)
  out = sq_eval.detections_to_all_boxes(recs, 0, None, 3)
  assert out == [[], [], []]
