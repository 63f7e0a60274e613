#!/usr/bin/env python
"""Computes the per-layer gains of the product's synthetic initialiser
(squeezedet_b200/utils/synth_gains.json) with the oracle's single-pass calibrator on
a small synthetic batch.  Run in the dev container:  python tests/golden/make_calibration.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from oracle import calib  # noqa: E402
from squeezedet_b200.utils import synth  # noqa: E402

SIZES = {'squeezeDet': (188, 624), 'squeezeDet+': (188, 624), 'vgg16': (96, 320),
         'resnet50': (188, 624)}


def main():
  out = {}
  for net, (h, w) in SIZES.items():
    specs = oracle.param_specs(net)
    raw = synth.synthetic_weights(specs, seed=0, gains={})
    x = synth.synthetic_images(2, h, w, seed=1234)
    _, gains = calib.calibrate(net, raw, x)
    out[net] = {k: round(float(v), 5) for k, v in gains.items()}
    print(net, len(gains), 'layers; gain range', min(gains.values()), max(gains.values()))
  with open(os.path.join(ROOT, 'squeezedet_b200', 'utils', 'synth_gains.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)


if __name__ == '__main__':
  main()
