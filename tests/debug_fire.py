"""GPU debug driver for the single-kernel fire module (fire_tc.cu): one shape per process so that
a hung kernel costs one `timeout`, with the error broken down by output half / tile position.

  python tests/debug_fire.py <shape index 0..9> [B H W]      # SqueezeDet fire2..fire11
Environment: SQDET_FUSED_FIRE=2 forces the fused kernel for every shape it takes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import oracle
from squeezedet_b200 import _lib
from gpu_util import fire_gpu, rel_err
from test_gpu_fire import SQUEEZEDET_FIRES, make_case, fire_oracle

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
spatial = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (2, 19, 37)
shape = SQUEEZEDET_FIRES[idx]
args = make_case(shape, spatial, seed=shape[0] * 7 + shape[1])
x, ws, bs, w1, b1, w3, b3 = args
bs = np.abs(bs) + 0.5        # bias-dominated squeeze: halo pixels outside the image must be 0
args = (x, ws, bs, w1, b1, w3, b3)
want = fire_oracle(*args, dtype=np.float64)
t0 = time.time()
got = fire_gpu(*args, math_mode=_lib.MATH_TF32X3_TC, device=0)
dt = time.time() - t0
E1 = w1.shape[3]
scale = np.abs(want).max()
err = np.abs(got.astype(np.float64) - want) / scale
nan = int(np.isnan(got).sum())
print('fire shape %s spatial %s: rel_err %.3e (e1 %.3e, e3 %.3e) nan %d  [%.2fs]' % (
    shape, spatial, np.nanmax(err) if nan < err.size else float('nan'),
    np.nanmax(err[..., :E1]), np.nanmax(err[..., E1:]), nan, dt))
if nan or np.nanmax(err) > 3e-5:
  e = np.where(np.isnan(err), 1.0, err)
  b, h, w, c = np.unravel_index(np.argmax(e), e.shape)
  print('  worst at (b %d, h %d, w %d, c %d): got %r want %r' % (b, h, w, c, got[b, h, w, c], want[b, h, w, c]))
  bad = e > 3e-5
  print('  bad fraction %.4f; by h: %s' % (bad.mean(), np.round(bad.mean(axis=(0, 2, 3)), 2).tolist()))
  print('  by w: %s' % np.round(bad.mean(axis=(0, 1, 3)), 2).tolist())
  cg = bad.mean(axis=(0, 1, 2)).reshape(-1, 16).mean(axis=1)
  print('  by 16-channel group: %s' % np.round(cg, 2).tolist())
  sys.exit(1)
