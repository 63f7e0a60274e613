"""End-to-end GPU parity through the reference-facing Python surface
(Net(mc, gpu_id) / Session.run / filter_prediction), oracle = numpy restatement."""
import numpy as np
import pytest

import oracle
from squeezedet_b200 import _lib, Session
from squeezedet_b200 import config as cfg
from squeezedet_b200.nets import SqueezeDet, SqueezeDetPlus, VGG16ConvDet, ResNet50ConvDet
from squeezedet_b200.utils import synth
from gpu_util import assert_classes_match, rel_err

pytestmark = pytest.mark.gpu

NETS = {
    'squeezeDet': (SqueezeDet, cfg.kitti_squeezeDet_config),
    'squeezeDet+': (SqueezeDetPlus, cfg.kitti_squeezeDetPlus_config),
    'vgg16': (VGG16ConvDet, cfg.kitti_vgg16_config),
    'resnet50': (ResNet50ConvDet, cfg.kitti_res50_config),
}
MODES = [_lib.MATH_FP32_SIMT, _lib.MATH_TF32X3_TC]

# Tolerances (BASELINE.json north_star): scores and box coordinates within 1e-4 relative;
# class ids / kept-box indices exact wherever the oracle's own margin exceeds fp noise.
TOL = 1e-4


def make_mc(net, width, height, batch):
  mc = NETS[net][1]()
  mc.IMAGE_WIDTH, mc.IMAGE_HEIGHT, mc.BATCH_SIZE = width, height, batch
  rows = oracle.layer_table(net, height, width)
  mc.GRID_H, mc.GRID_W = rows[-1][2][0], rows[-1][2][1]
  mc.ANCHOR_BOX = cfg.set_anchors(mc)
  mc.ANCHORS = len(mc.ANCHOR_BOX)
  return mc


def assert_boxes_close(got, ref32, ref64):
  """Box coordinates: within 1e-4 relative of the fp32 reference, plus the reference's OWN
  fp32 uncertainty (|ref32 - ref64|, x4) — boxes that clip from ~4000 px wide pre-clip values
  carry ~1e-3 px of fp32 rounding in any implementation — plus 4e-3 px absolute (3e-6 of the
  image width)."""
  got = np.asarray(got, np.float64)
  tol = TOL * np.abs(ref32) + 4.0 * np.abs(np.asarray(ref32, np.float64) - ref64) + 4e-3
  bad = np.abs(got - ref32) > tol
  assert not bad.any(), (int(bad.sum()), float(np.abs(got - ref32)[bad].max()))


def oracle_run(net, mc, weights, images, dtype=np.float32, keep=None):
  preds = oracle.forward(net, weights, images, dtype=dtype, keep=keep)
  return preds, oracle.interpret_output(preds, mc.ANCHOR_BOX, mc.CLASSES, mc.ANCHOR_PER_GRID,
                                        mc.IMAGE_WIDTH, mc.IMAGE_HEIGHT, mc.EXP_THRESH, dtype)


@pytest.mark.parametrize('math_mode', MODES)
@pytest.mark.parametrize('net,width,height', [
    ('squeezeDet', 208, 112), ('squeezeDet+', 215, 119), ('vgg16', 96, 64),
    ('resnet50', 131, 99)])
def test_layerwise_parity_small_image(net, width, height, math_mode, gpu_device):
  mc = make_mc(net, width, height, 2)
  model = NETS[net][0](mc, gpu_device, math_mode=math_mode)
  weights = synth.synthetic_weights(synth.model_param_specs(model), seed=3)
  assert [n for n, _ in synth.model_param_specs(model)] == [n for n, _ in oracle.param_specs(net)]
  model.load_weights(weights)
  images = synth.synthetic_images(2, height, width, seed=9)
  keep64, keep32 = {}, {}
  p64, (b64, s64, c64) = oracle_run(net, mc, weights, images, np.float64, keep64)
  p32, (b32, s32, c32) = oracle_run(net, mc, weights, images, np.float32, keep32)
  boxes, probs, cls = model.detect(images)
  checked = 0
  for name, want in keep64.items():
    if name in model._tensors:
      try:
        got = model.read_tensor(name)
      except _lib.SqdetError as exc:
        assert exc.code == -5, exc          # fused away (e.g. fire3 when pool3 is fused)
        continue
      assert got.shape == want.shape, name
      # (1) the bar: within 1e-4 (relative to the tensor's scale) of the fp32 reference
      assert rel_err(got, keep32[name]) < TOL, (name, rel_err(got, keep32[name]))
      # (2) quality: as close to the fp64 truth as the fp32 reference itself is (x4 slack)
      e_gpu, e_ref = rel_err(got, want), rel_err(keep32[name], want)
      assert e_gpu < max(4 * e_ref, 2e-5), (name, e_gpu, e_ref)
      checked += 1
  assert checked >= 10
  np.testing.assert_allclose(probs, s32, rtol=TOL, atol=1e-7)
  assert_boxes_close(boxes, b32, b64)
  assert_classes_match(cls, c32, p64, mc.ANCHOR_PER_GRID, mc.CLASSES, TOL)


@pytest.mark.parametrize('math_mode', MODES)
def test_full_size_squeezedet_detections(math_mode, gpu_device):
  """SqueezeDet at the BASELINE config size (1242x375), b=2: det tensors within 1e-4 of
  the fp32 oracle; filtered records identical to running the oracle's filter_prediction on
  the oracle's det tensors wherever the oracle's top-65 score gaps exceed the tolerance."""
  net = 'squeezeDet'
  mc = make_mc(net, 1242, 375, 2)
  assert (mc.GRID_H, mc.GRID_W, mc.ANCHORS) == (24, 78, 16848)
  model = SqueezeDet(mc, gpu_device, math_mode=math_mode)
  weights = synth.synthetic_weights(synth.model_param_specs(model), seed=0)
  model.load_weights(weights)
  images = synth.synthetic_images(2, 375, 1242, seed=1234)
  _, (wb, wp, wc) = oracle_run(net, mc, weights, images, np.float32)
  p64, (wb64, _, _) = oracle_run(net, mc, weights, images, np.float64)
  boxes, probs, cls, dets, counts = model.detect(images, want_dets=True)
  assert boxes.dtype == np.float32 and probs.dtype == np.float32 and cls.dtype == np.int64
  np.testing.assert_allclose(probs, wp, rtol=TOL, atol=1e-7)
  assert_boxes_close(boxes, wb, wb64)
  assert_classes_match(cls, wc, p64, mc.ANCHOR_PER_GRID, mc.CLASSES, TOL)
  for i in range(2):
    # (1) bit-exact: GPU filter on the GPU's own det tensors == oracle filter on them
    fb, fp, fc, src = oracle.filter_prediction(boxes[i], probs[i], cls[i], mc.CLASSES,
                                               mc.TOP_N_DETECTION, mc.PROB_THRESH, mc.NMS_THRESH)
    n = int(counts[i])
    assert n == len(src)
    assert dets[i]['anchor'][:n].tolist() == src
    assert dets[i]['cls'][:n].tolist() == fc
    assert np.array_equal(dets[i]['prob'][:n], np.asarray(fp, np.float32))
    # (2) margin-aware vs the oracle's own pipeline: kept-box indices must agree except for
    # anchors whose oracle score sits within 10*TOL of another top-66 score (a near tie whose
    # order fp noise may legitimately flip)
    ob, op, oc, osrc = oracle.filter_prediction(wb[i], wp[i], wc[i], mc.CLASSES,
                                                mc.TOP_N_DETECTION, mc.PROB_THRESH, mc.NMS_THRESH)
    order = np.argsort(-wp[i].astype(np.float64), kind='stable')[:66]
    top = wp[i][order].astype(np.float64)
    near_tie = set()
    for a in range(len(top)):
      for b in range(len(top)):
        if a != b and abs(top[a] - top[b]) <= 10 * TOL * top[a]:
          near_tie.add(int(order[a]))
    diff = set(src) ^ set(osrc)
    assert diff <= near_tie, (sorted(diff), sorted(near_tie))
    if not near_tie:
      assert src == osrc


@pytest.mark.parametrize('math_mode', MODES)
def test_session_run_contract_and_filter_prediction(math_mode, gpu_device):
  """The reference call shape: sess.run([det_boxes, det_probs, det_class], feed_dict) then
  model.filter_prediction per image (demo.py:193-199)."""
  mc = make_mc('squeezeDet', 416, 128, 1)
  model = SqueezeDet(mc, gpu_device, math_mode=math_mode)
  model.load_weights(synth.synthetic_weights(synth.model_param_specs(model), seed=5))
  img = synth.synthetic_images(1, 128, 416, seed=6)[0]
  with Session() as sess:
    det_boxes, det_probs, det_class = sess.run(
        [model.det_boxes, model.det_probs, model.det_class],
        feed_dict={model.image_input: [img]})
  A = mc.ANCHORS
  assert det_boxes.shape == (1, A, 4) and det_probs.shape == (1, A) and det_class.shape == (1, A)
  final_boxes, final_probs, final_class = model.filter_prediction(
      det_boxes[0], det_probs[0], det_class[0])
  ob, op, oc, _ = oracle.filter_prediction(det_boxes[0], det_probs[0], det_class[0], mc.CLASSES,
                                           mc.TOP_N_DETECTION, mc.PROB_THRESH, mc.NMS_THRESH)
  assert final_class == oc
  assert all(np.array_equal(a, b) for a, b in zip(final_boxes, ob))
  assert [float(x) for x in final_probs] == [float(x) for x in op]
  assert isinstance(final_boxes, list) and isinstance(final_class[0], int)
  # one-pass variant gives the same triple
  fb2, fp2, fc2 = model.detect_filtered([img])[0]
  assert fc2 == final_class and all(np.array_equal(a, b) for a, b in zip(fb2, final_boxes))
  # static feed shape, like the TF placeholder
  with pytest.raises(ValueError):
    model.detect(np.zeros((2, 128, 416, 3), np.float32))
  # counters follow the reference formulas (nn_skeleton.py:549-561)
  assert sum(v for _, v in model.model_size_counter) == 2082120
  assert len(model.model_params) == 64


def test_batch_invariance_and_determinism(gpu_device):
  mc1 = make_mc('squeezeDet', 320, 96, 1)
  mc3 = make_mc('squeezeDet', 320, 96, 3)
  m1 = SqueezeDet(mc1, gpu_device)
  m3 = SqueezeDet(mc3, gpu_device)
  w = synth.synthetic_weights(synth.model_param_specs(m1), seed=8)
  m1.load_weights(w)
  m3.load_weights(w)
  imgs = synth.synthetic_images(3, 96, 320, seed=4)
  b3, p3, c3 = m3.detect(imgs)
  b3b, p3b, c3b = m3.detect(imgs)
  assert np.array_equal(p3, p3b) and np.array_equal(b3, b3b) and np.array_equal(c3, c3b)
  for i in range(3):
    b1, p1, c1 = m1.detect(imgs[i:i + 1])
    assert np.array_equal(p1[0], p3[i]) and np.array_equal(b1[0], b3[i])


def test_set_param_errors(gpu_device):
  mc = make_mc('squeezeDet', 160, 96, 1)
  m = SqueezeDet(mc, gpu_device)
  with pytest.raises(_lib.SqdetError):
    m.set_param('conv1/kernels', np.zeros((3, 3, 3, 63), np.float32))
  with pytest.raises(_lib.SqdetError):
    m.set_param('nope/kernels', np.zeros((1,), np.float32))


def test_pipelined_submit_and_uint8_input(gpu_device):
  """sqdet_submit/sqdet_wait (depth-2 pipeline) and the uint8 path: same records as the
  synchronous fp32 feed of `im - BGR_MEANS` (demo.py:187-190)."""
  mc = make_mc('squeezeDet', 320, 96, 2)
  m = SqueezeDet(mc, gpu_device)
  m.load_weights(synth.synthetic_weights(synth.model_param_specs(m), seed=8))
  rng = np.random.default_rng(3)
  batches = [rng.integers(0, 256, (2, 96, 320, 3), dtype=np.uint8) for _ in range(4)]
  feeds = [(b.astype(np.float32) - np.asarray(mc.BGR_MEANS)).astype(np.float32) for b in batches]
  want = [m.detect_records(f) for f in feeds]
  # uint8 path, synchronous convenience
  for b, (wd, wc) in zip(batches, want):
    d, c = m.detect_u8(b)
    assert np.array_equal(c, wc) and np.array_equal(d, wd)
  # pipelined, two in flight, mixing input types
  outs = [(np.empty((2, m.max_dets), _lib.DET_DTYPE), np.empty((2,), np.int32)) for _ in batches]
  keep_alive = []
  for i in range(len(batches)):
    src = np.ascontiguousarray(batches[i]) if i % 2 == 0 else np.ascontiguousarray(feeds[i])
    keep_alive.append(src)
    m.submit(src.ctypes.data, outs[i][0].ctypes.data, outs[i][1].ctypes.data,
             _lib.IMG_U8 if i % 2 == 0 else _lib.IMG_F32)
    if i >= 1:
      m.wait()
  m.wait()
  for (d, c), (wd, wc) in zip(outs, want):
    assert np.array_equal(c, wc) and np.array_equal(d, wd)
  with pytest.raises(_lib.SqdetError):
    m.wait()                                   # nothing in flight
