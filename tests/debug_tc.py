"""GPU debug driver for the tcgen05 conv path: progressively harder cases, one line each,
flushed, so a hang or a layout bug is attributable.  Run under `timeout`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle
from gpu_util import conv2d_gpu

def run(tag, B, H, W, Cin, Cout, k, relu=True, seed=0, structured=None, y_cstride=None, y_coff=0):
  rng = np.random.default_rng(seed)
  if structured == 'delta':
    x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
    w = np.zeros((k, k, Cin, Cout), np.float32)
    for c in range(min(Cin, Cout)):
      w[k // 2, k // 2, c, c] = 1.0
    b = np.zeros(Cout, np.float32)
  else:
    x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.normal(size=(Cout,)).astype(np.float32)
  want = oracle.conv2d(x, w, b, 1, 'SAME', relu, np.float64)
  print(tag, '...', end=' ', flush=True)
  got = conv2d_gpu(x, w, b, 1, 'SAME', relu=relu, math_mode=1, y_cstride=y_cstride, y_coff=y_coff)
  if y_cstride:
    got = got[..., y_coff:y_coff + Cout]
  err = np.abs(got - want).max() / np.abs(want).max()
  simt = conv2d_gpu(x, w, b, 1, 'SAME', relu=relu, math_mode=0)
  err_s = np.abs(simt - want).max() / np.abs(want).max()
  print('rel err tc %.3e  simt %.3e  nan=%d' % (err, err_s, int(np.isnan(got).sum())), flush=True)
  if err > 1e-4:
    bad = np.argwhere(np.abs(got - want) > 1e-3 * np.abs(want).max())
    print('   first bad idx', bad[:6].tolist(), 'n_bad', len(bad), 'of', got.size, flush=True)
    print('   got', got.reshape(-1)[:8], '\n   want', want.reshape(-1)[:8], flush=True)
  return err

cases = [
  ('1x1 delta Cin32 Cout32 1tile', dict(B=1, H=8, W=16, Cin=32, Cout=32, k=1, relu=False, structured='delta')),
  ('1x1 rand  Cin32 Cout32 1tile', dict(B=1, H=8, W=16, Cin=32, Cout=32, k=1)),
  ('1x1 rand  Cin16 Cout16 (SW64)', dict(B=1, H=8, W=16, Cin=16, Cout=16, k=1)),
  ('1x1 rand  Cin64 Cout64 kch2', dict(B=1, H=8, W=16, Cin=64, Cout=64, k=1)),
  ('3x3 delta Cin32 Cout32', dict(B=1, H=8, W=16, Cin=32, Cout=32, k=3, relu=False, structured='delta')),
  ('3x3 rand  Cin32 Cout32', dict(B=1, H=8, W=16, Cin=32, Cout=32, k=3)),
  ('3x3 rand  ragged 13x29 Cin48 Cout192', dict(B=2, H=13, W=29, Cin=48, Cout=192, k=3)),
  ('3x3 rand  Cin16 Cout64 24x31', dict(B=2, H=24, W=31, Cin=16, Cout=64, k=3)),
  ('1x1 rand  Cin96 Cout384 (2 chunks)', dict(B=1, H=12, W=20, Cin=96, Cout=384, k=1)),
  ('3x3 rand  Cin256 Cout72 (N=80)', dict(B=1, H=12, W=20, Cin=256, Cout=72, k=3, relu=False)),
  ('3x3 rand  Cin768 Cout72 head K=6912', dict(B=1, H=24, W=78, Cin=768, Cout=72, k=3, relu=False)),
  ('3x3 rand  window 80/16', dict(B=1, H=15, W=18, Cin=32, Cout=48, k=3, y_cstride=80, y_coff=16)),
  ('1x1 rand  Cin512 Cout64 big', dict(B=4, H=47, W=156, Cin=512, Cout=64, k=1)),
]
only = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
for i, (tag, kw) in enumerate(cases):
  if only and i not in only:
    continue
  run('[%d] %s' % (i, tag), **kw)
print('done', flush=True)
