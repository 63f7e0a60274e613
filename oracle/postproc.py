"""Anchors, interpret_output and filter_prediction/NMS restated in numpy
(oracle — test infrastructure only).

Follows (reference, read-only):
  * ``src/config/kitti_squeezeDet_config.py:45-79``  ``set_anchors``
  * ``src/nn_skeleton.py:142-238,271-283``           ``_add_interpretation_graph``
  * ``src/utils/util.py:167-196``                    ``bbox_transform[_inv]``
  * ``src/utils/util.py:219-231``                    ``safe_exp``
  * ``src/utils/util.py:32-54,56-76``                ``batch_iou``, ``nms``
  * ``src/nn_skeleton.py:696-734``                   ``filter_prediction``
``batch_iou`` / ``nms`` / ``filter_prediction`` / ``set_anchors`` are PINNED
against the reference's own functions (tests/test_oracle_pinning.py and the
fixtures made by tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

# The nine anchor shapes (w, h) per grid cell.
#   kitti_squeezeDet_config.py:49-51 (also squeezeDet+ and vgg16 configs)
ANCHOR_SHAPES_SQUEEZE = np.array(
    [[36., 37.], [366., 174.], [115., 59.], [162., 87.], [38., 90.],
     [258., 173.], [224., 108.], [78., 170.], [72., 43.]])
#   kitti_res50_config.py:49-51
ANCHOR_SHAPES_RES50 = np.array(
    [[94., 49.], [225., 161.], [170., 91.], [390., 181.], [41., 32.],
     [128., 64.], [298., 164.], [232., 99.], [65., 42.]])


def set_anchors(image_width, image_height, grid_h, grid_w, shapes):
  """[A,4] float64 (cx, cy, w, h), A = grid_h*grid_w*len(shapes), ordered
  (row i, col j, shape k) — anchor id = (i*grid_w + j)*K + k.

  cx_j = (j+1) * float(W_img) / (grid_w+1)  (multiply first, then divide, in
  float64 — kitti_squeezeDet_config.py:57), cy_i likewise (:67).
  """
  shapes = np.asarray(shapes, dtype=np.float64)
  K = shapes.shape[0]
  cx = np.arange(1, grid_w + 1) * float(image_width) / (grid_w + 1)
  cy = np.arange(1, grid_h + 1) * float(image_height) / (grid_h + 1)
  out = np.empty((grid_h, grid_w, K, 4), dtype=np.float64)
  out[..., 0] = cx[None, :, None]
  out[..., 1] = cy[:, None, None]
  out[..., 2] = shapes[None, None, :, 0]
  out[..., 3] = shapes[None, None, :, 1]
  return out.reshape(-1, 4)


def safe_exp(w, thresh, dtype=np.float32):
  """util.py:219-231.  slope = np.exp(thresh) (float64, cast when it meets the
  fp32 tensor); out = lin*(slope*(w-thresh+1)) + (1-lin)*exp(where(w>thresh,0,w))."""
  w = np.asarray(w, dtype=dtype)
  slope = dtype(np.exp(thresh))
  t = dtype(thresh)
  one = dtype(1.0)
  lin_bool = w > t
  lin = lin_bool.astype(dtype)
  lin_out = slope * (w - t + one)
  exp_out = np.exp(np.where(lin_bool, dtype(0), w))
  return lin * lin_out + (one - lin) * exp_out


def interpret_output(preds, anchors, classes, anchors_per_grid, image_width,
                     image_height, exp_thresh=1.0, dtype=np.float32):
  """preds [B,Hg,Wg,K*(C+1+4)] -> det_boxes [B,A,4] (cx,cy,w,h), det_probs
  [B,A], det_class [B,A] int64   (nn_skeleton.py:146-238, 271-283).

  Channel layout: [0,K*C) class logits (k*C+c), then K
  confidence logits, then K*4 deltas (k*4 + {dx,dy,dw,dh}).
  """
  preds = np.asarray(preds, dtype=dtype)
  B = preds.shape[0]
  K, C = anchors_per_grid, classes
  A = preds.shape[1] * preds.shape[2] * K
  anc = np.asarray(anchors, dtype=np.float64).astype(dtype)
  assert anc.shape == (A, 4), (anc.shape, A)

  # class probabilities: softmax over C (tf.nn.softmax, max-subtracted)
  logits = preds[..., :K * C].reshape(-1, C)
  z = logits - logits.max(axis=1, keepdims=True)
  e = np.exp(z)
  class_probs = (e / e.sum(axis=1, keepdims=True)).reshape(B, A, C)
  # confidence: tf.sigmoid
  conf_logit = preds[..., K * C:K * C + K].reshape(B, A)
  one = dtype(1.0)
  conf = one / (one + np.exp(-conf_logit))
  delta = preds[..., K * C + K:].reshape(B, A, 4)

  ax, ay, aw, ah = anc[:, 0], anc[:, 1], anc[:, 2], anc[:, 3]
  cx = ax + delta[..., 0] * aw
  cy = ay + delta[..., 1] * ah
  bw = aw * safe_exp(delta[..., 2], exp_thresh, dtype)
  bh = ah * safe_exp(delta[..., 3], exp_thresh, dtype)

  two = dtype(2.0)
  xmin, ymin = cx - bw / two, cy - bh / two      # util.py:174-177
  xmax, ymax = cx + bw / two, cy + bh / two
  wm1, hm1 = dtype(image_width - 1.0), dtype(image_height - 1.0)
  zero = dtype(0.0)
  xmin = np.minimum(np.maximum(zero, xmin), wm1)  # nn_skeleton.py:219-233
  ymin = np.minimum(np.maximum(zero, ymin), hm1)
  xmax = np.maximum(np.minimum(wm1, xmax), zero)
  ymax = np.maximum(np.minimum(hm1, ymax), zero)
  width = xmax - xmin + one                       # util.py:189-194
  height = ymax - ymin + one
  half = dtype(0.5)
  det_boxes = np.stack(
      [xmin + half * width, ymin + half * height, width, height], axis=-1)

  probs = class_probs * conf[..., None]           # nn_skeleton.py:274-278
  det_probs = probs.max(axis=2)
  det_class = probs.argmax(axis=2).astype(np.int64)
  return det_boxes, det_probs, det_class


def batch_iou(boxes, box):
  """util.py:32-54 — centre-format IoU of each row of `boxes` against `box`,
  evaluated in the arrays' own dtype, in the reference's operation order."""
  lr = np.maximum(
      np.minimum(boxes[:, 0] + 0.5 * boxes[:, 2], box[0] + 0.5 * box[2]) -
      np.maximum(boxes[:, 0] - 0.5 * boxes[:, 2], box[0] - 0.5 * box[2]), 0)
  tb = np.maximum(
      np.minimum(boxes[:, 1] + 0.5 * boxes[:, 3], box[1] + 0.5 * box[3]) -
      np.maximum(boxes[:, 1] - 0.5 * boxes[:, 3], box[1] - 0.5 * box[3]), 0)
  inter = lr * tb
  union = boxes[:, 2] * boxes[:, 3] + box[2] * box[3] - inter
  with np.errstate(divide='ignore', invalid='ignore'):
    return inter / union


def _rank_order(probs):
  """Descending order, ties broken by ascending index (the engine's documented
  rule; the reference's ``argsort()[::-1]`` leaves tie order unspecified)."""
  probs = np.asarray(probs)
  return np.lexsort((np.arange(len(probs)), -probs.astype(np.float64)))


def nms(boxes, probs, threshold):
  """util.py:56-76 — NOT greedy NMS: a box that is itself suppressed still
  suppresses lower-ranked boxes, i.e.
      keep[j] = not exists i ranked above j with IoU(i, j) > threshold.
  The comparison follows the installed numpy (NEP 50): float32 IoU against
  float32(threshold)."""
  boxes = np.asarray(boxes)
  probs = np.asarray(probs)
  n = len(probs)
  order = _rank_order(probs)
  keep = [True] * n
  thr = boxes.dtype.type(threshold) if n else threshold
  for i in range(n - 1):
    ovps = batch_iou(boxes[order[i + 1:]], boxes[order[i]])
    for j in np.nonzero(ovps > thr)[0]:
      keep[order[j + i + 1]] = False
  return keep


def filter_prediction(boxes, probs, cls_idx, classes, top_n, prob_thresh,
                      nms_thresh):
  """nn_skeleton.py:696-734.  Returns (final_boxes, final_probs, final_cls,
  final_src) — the first three as the reference returns them (lists grouped
  by class id ascending, inside a class in descending-prob order); `final_src`
  additionally gives each kept box's index into the input arrays."""
  boxes = np.asarray(boxes)
  probs = np.asarray(probs)
  cls_idx = np.asarray(cls_idx)
  if 0 < top_n < len(probs):
    order = _rank_order(probs)[:top_n]            # :711-715
  else:
    order = np.nonzero(probs > prob_thresh)[0]    # :716-720 (original order)
  probs, boxes, cls_idx = probs[order], boxes[order], cls_idx[order]
  final_boxes, final_probs, final_cls, final_src = [], [], [], []
  for c in range(classes):
    idx_c = [i for i in range(len(probs)) if cls_idx[i] == c]
    keep = nms(boxes[idx_c], probs[idx_c], nms_thresh)
    for i, k in enumerate(keep):
      if k:
        final_boxes.append(boxes[idx_c[i]])
        final_probs.append(probs[idx_c[i]])
        final_cls.append(c)
        final_src.append(int(order[idx_c[i]]))
  return final_boxes, final_probs, final_cls, final_src
