"""Drawing and KITTI detection-file helpers used by the entry points (host side, "next"
rows f-3 / f-4 of SURVEY.md §8)."""
from __future__ import annotations

import os

import numpy as np

from .util import bbox_transform

# BGR colours per class name, as the reference demo uses (src/demo.py:207-212)
CLASS_COLORS = {'car': (255, 191, 0), 'cyclist': (0, 191, 255), 'pedestrian': (255, 0, 191)}


def draw_box(im, box_list, label_list, color=(0, 255, 0), cdict=None, form='center'):
  """Rectangle + 'CLASS: (PROB)' label per detection, in place (reference
  src/train.py:51-72 `_draw_box`: 1-px rectangle, FONT_HERSHEY_SIMPLEX 0.3 at (xmin, ymax))."""
  import cv2
  if form not in ('center', 'diagonal'):
    raise ValueError('bounding box format not accepted: {}.'.format(form))
  for bbox, label in zip(box_list, label_list):
    if form == 'center':
      bbox = bbox_transform(bbox)
    xmin, ymin, xmax, ymax = [int(b) for b in bbox]
    name = label.split(':')[0]
    c = cdict[name] if cdict and name in cdict else color
    cv2.rectangle(im, (xmin, ymin), (xmax, ymax), c, 1)
    cv2.putText(im, label, (xmin, ymax), cv2.FONT_HERSHEY_SIMPLEX, 0.3, c, 1)
  return im


def kitti_detection_line(cls_name, box_xyxy, score):
  """One line of a KITTI 2-D detection file (reference src/dataset/kitti.py:116-127):
  type, truncated -1, occluded -1, alpha 0.0, bbox with 2 decimals, 7 zeros, score %.3f."""
  return ('{:s} -1 -1 0.0 {:.2f} {:.2f} {:.2f} {:.2f} 0.0 0.0 0.0 0.0 0.0 0.0 0.0 {:.3f}\n'
          .format(cls_name.lower(), box_xyxy[0], box_xyxy[1], box_xyxy[2], box_xyxy[3], score))


def write_kitti_detections(det_file_dir, image_ids, class_names, all_boxes):
  """all_boxes[cls][image] = list/array of [xmin, ymin, xmax, ymax, score]
  (reference kitti.py:100-127).  Returns the directory that holds `data/`."""
  os.makedirs(det_file_dir, exist_ok=True)
  for im_idx, index in enumerate(image_ids):
    with open(os.path.join(det_file_dir, index + '.txt'), 'wt') as f:
      for cls_idx, cls in enumerate(class_names):
        for det in all_boxes[cls_idx][im_idx]:
          f.write(kitti_detection_line(cls, det[:4], det[4]))
  return os.path.dirname(det_file_dir)


def parse_kitti_ap_files(result_dir, class_names):
  """stats_<cls>_ap.txt written by the (unmodified) evaluate_object binary
  (reference kitti.py:138-159) -> (aps, names)."""
  aps, names = [], []
  for cls in class_names:
    path = os.path.join(result_dir, 'stats_{:s}_ap.txt'.format(cls))
    if os.path.exists(path):
      with open(path) as f:
        lines = f.readlines()
      assert len(lines) == 3, 'Line number of {} should be 3'.format(path)
      aps += [float(line.split('=')[1].strip()) for line in lines]
    else:
      aps += [0.0, 0.0, 0.0]
    names += [cls + '_easy', cls + '_medium', cls + '_hard']
  return aps, names
