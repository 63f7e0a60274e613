#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ by RUNNING THE
REFERENCE'S OWN CODE, imported unmodified from /root/reference/src (dev container
only; see oracle/ref_import.py for how the py2/TF imports are stubbed).

  python tests/golden/make_golden.py

Outputs
  postproc_kat.npz   filter_prediction / nms / batch_iou known-answer cases:
                     inputs + the reference's outputs (src/nn_skeleton.py:696-734,
                     src/utils/util.py:32-76)
  anchors.json       sha256 + corner rows of mc.ANCHOR_BOX for the four KITTI configs
                     (src/config/kitti_*_config.py set_anchors), plus the scalar mc keys
The TensorFlow half of the path cannot be run (no TF here), so there are no
reference-generated conv vectors: that half stays "parity unpinned".
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402


def rand_case(rng, n, img_w=1242., img_h=375., classes=3, clustered=False):
  if clustered:   # many overlapping boxes around a few centres -> NMS does real work
    k = max(1, n // 12)
    centres = np.stack([rng.uniform(50, img_w - 50, k), rng.uniform(30, img_h - 30, k)], 1)
    pick = rng.integers(0, k, n)
    cx = centres[pick, 0] + rng.normal(0, 12, n)
    cy = centres[pick, 1] + rng.normal(0, 8, n)
    w = rng.uniform(40, 160, n)
    h = rng.uniform(30, 120, n)
  else:
    cx, cy = rng.uniform(0, img_w, n), rng.uniform(0, img_h, n)
    w, h = rng.uniform(4, 300, n), rng.uniform(4, 200, n)
  boxes = np.stack([cx, cy, w, h], 1).astype(np.float32)
  probs = rng.uniform(0, 1, n).astype(np.float32)
  # make scores distinct so the unspecified argsort tie order never matters
  probs = np.unique(probs)
  while len(probs) < n:
    probs = np.unique(np.concatenate([probs, rng.uniform(0, 1, n).astype(np.float32)]))
  probs = rng.permutation(probs)[:n].astype(np.float32)
  cls = rng.integers(0, classes, n).astype(np.int64)
  return boxes, probs, cls


def main():
  ns = ref_import.load()
  rng = np.random.default_rng(20260922)
  cases = {}
  meta = []

  def add(name, boxes, probs, cls, classes, top_n, prob_thresh, nms_thresh):
    fb, fp, fc = ref_import.ref_filter_prediction(ns, boxes, probs, cls, classes, top_n,
                                                  prob_thresh, nms_thresh)
    i = len(meta)
    cases['c%d_boxes' % i] = boxes
    cases['c%d_probs' % i] = probs
    cases['c%d_cls' % i] = cls
    cases['c%d_out_boxes' % i] = np.asarray(fb, np.float32).reshape(-1, 4)
    cases['c%d_out_probs' % i] = np.asarray(fp, np.float32)
    cases['c%d_out_cls' % i] = np.asarray(fc, np.int64)
    cases['c%d_nms_keep' % i] = np.asarray(
        ns.util.nms(boxes, probs, nms_thresh), bool) if len(probs) <= 512 else np.zeros(0, bool)
    meta.append(dict(name=name, classes=classes, top_n=top_n, prob_thresh=prob_thresh,
                     nms_thresh=nms_thresh, n=int(len(probs))))

  # 1. random and clustered, top-N branch (the configured path: TOP_N_DETECTION=64)
  for n, clustered in ((200, False), (200, True), (1000, True), (65, True), (4000, True)):
    b, p, c = rand_case(rng, n, clustered=clustered)
    add('topn64_n%d_%s' % (n, 'clustered' if clustered else 'uniform'), b, p, c, 3, 64,
        0.005, 0.4)
  # 2. full anchor count (16848) like one real image
  b, p, c = rand_case(rng, 16848, clustered=True)
  add('topn64_full_16848', b, p, c, 3, 64, 0.005, 0.4)
  # 3. threshold branch: n <= TOP_N_DETECTION, and TOP_N_DETECTION == 0
  b, p, c = rand_case(rng, 64, clustered=True)
  add('thresh_n_eq_topn', b, p, c, 3, 64, 0.005, 0.4)
  b, p, c = rand_case(rng, 40, clustered=True)
  p[:7] = np.float32(0.001) * np.arange(1, 8, dtype=np.float32) / 8     # below PROB_THRESH
  add('thresh_small_with_lowprob', b, p, c, 3, 64, 0.005, 0.4)
  b, p, c = rand_case(rng, 300, clustered=True)
  add('thresh_topn0', b, p, c, 3, 0, 0.5, 0.4)
  # 4. the chain A>B>C where a suppressed box still suppresses (NOT greedy NMS)
  b = np.array([[50, 50, 100, 100], [60, 50, 100, 100], [90, 50, 100, 100]], np.float32)
  p = np.array([.9, .8, .7], np.float32)
  add('chain_suppressed_still_suppresses', b, p, np.zeros(3, np.int64), 1, 64, 0.005, 0.3)
  # 5. IoU exactly at / just around the threshold: boxes built so inter/union == 0.4 in fp32
  b = np.array([[100, 100, 70, 10], [130, 100, 70, 10], [300, 100, 70, 10],
                [330.001, 100, 70, 10]], np.float32)   # 40/100 overlap -> IoU 0.4
  p = np.array([.9, .8, .7, .6], np.float32)
  add('iou_at_threshold', b, p, np.zeros(4, np.int64), 1, 64, 0.005, 0.4)
  # 6. degenerate boxes (zero area -> 0/0) and many classes (PASCAL_VOC = 20)
  b, p, c = rand_case(rng, 120, clustered=True, classes=20)
  b[:5, 2:] = 0
  add('voc20_with_zero_area', b, p, c, 20, 64, 0.005, 0.2)
  # 7. single box, single class
  add('single', np.array([[10, 10, 5, 5]], np.float32), np.array([.5], np.float32),
      np.zeros(1, np.int64), 3, 64, 0.005, 0.4)
  # 8. class ids outside range(classes) are silently dropped by the reference loop
  b, p, c = rand_case(rng, 100, clustered=True, classes=5)
  add('out_of_range_classes', b, p, c, 3, 64, 0.005, 0.4)

  # batch_iou direct KAT
  b, _, _ = rand_case(rng, 257, clustered=True)
  cases['iou_boxes'] = b
  cases['iou_out'] = np.stack([ns.util.batch_iou(b, b[i]) for i in (0, 17, 256)])

  cases['meta_json'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
  np.savez_compressed(os.path.join(HERE, 'postproc_kat.npz'), **cases)

  anchors = {}
  for fn in sorted(ns.configs):
    mc = ns.configs[fn]()
    ab = np.ascontiguousarray(mc.ANCHOR_BOX, dtype=np.float64)
    scal = {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item())
            for k, v in mc.items()
            if isinstance(v, (int, float, str, bool, np.floating, np.integer))}
    anchors[fn] = dict(sha256=hashlib.sha256(ab.tobytes()).hexdigest(), shape=list(ab.shape),
                       first=ab[0].tolist(), last=ab[-1].tolist(), row_1000=ab[1000].tolist(),
                       class_names=list(mc.CLASS_NAMES),
                       bgr_means=np.asarray(mc.BGR_MEANS).ravel().tolist(), scalars=scal)
  with open(os.path.join(HERE, 'anchors.json'), 'w') as f:
    json.dump(anchors, f, indent=1, sort_keys=True)
  print('wrote', len(meta), 'post-proc cases and', len(anchors), 'configs')


if __name__ == '__main__':
  main()
