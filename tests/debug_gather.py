"""GPU debug driver for the first-layer (gather-mode) tensor-core conv: 3x3 over a 3-channel
image, stride 1 / 2, SAME / VALID, checked against the oracle.  Run under `timeout`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle
from gpu_util import conv2d_gpu

def run(tag, B, H, W, Cout, stride, padding, relu=True, seed=0):
  rng = np.random.default_rng(seed)
  x = (rng.normal(size=(B, H, W, 3)) * 60).astype(np.float32)
  w = (rng.normal(size=(3, 3, 3, Cout)) / np.sqrt(27)).astype(np.float32)
  b = rng.normal(size=(Cout,)).astype(np.float32)
  want = oracle.conv2d(x, w, b, stride, padding, relu, np.float64)
  print(tag, '...', end=' ', flush=True)
  got = conv2d_gpu(x, w, b, stride, padding, relu=relu, math_mode=1)
  simt = conv2d_gpu(x, w, b, stride, padding, relu=relu, math_mode=0)
  err = np.abs(got - want).max() / np.abs(want).max()
  err_s = np.abs(simt - want).max() / np.abs(want).max()
  print('shape', got.shape, 'rel err tc %.3e  simt %.3e  nan=%d' % (err, err_s, int(np.isnan(got).sum())), flush=True)
  if err > 1e-4:
    bad = np.argwhere(np.abs(got - want) > 1e-3 * np.abs(want).max())
    print('   first bad idx', bad[:6].tolist(), 'n_bad', len(bad), 'of', got.size, flush=True)
    print('   got', got.reshape(-1)[:8], '\n   want', want.reshape(-1)[:8], flush=True)

run('s1 SAME  8x16   C32', 1, 8, 16, 32, 1, 'SAME', relu=False)
run('s1 SAME  13x29  C64', 2, 13, 29, 64, 1, 'SAME')
run('s2 VALID 33x65  C64', 1, 33, 65, 64, 2, 'VALID')
run('s2 SAME  34x70  C64', 2, 34, 70, 64, 2, 'SAME')
run('s2 VALID 375x1242 C64', 2, 375, 1242, 64, 2, 'VALID')
run('s1 SAME  48x100 C96', 1, 48, 100, 96, 1, 'SAME')
print('done')
