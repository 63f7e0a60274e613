"""Worker for test_gloo_world2_allgather_roundtrip (CPU, gloo)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from squeezedet_b200 import shard
from squeezedet_b200._lib import DET_DTYPE

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
GLOBAL_B, MAX_DETS = 5, 8
counts_all = [2, 0, 5, 1, 3]
lo, hi = shard.shard_ranges(GLOBAL_B, world)[rank]
bmax = max(shard.shard_sizes(GLOBAL_B, world))
dets = np.zeros((bmax, MAX_DETS), DET_DTYPE)
dets['anchor'] = -1
counts = np.zeros((bmax,), np.int32)
for i, g in enumerate(range(lo, hi)):
  n = counts_all[g]
  counts[i] = n
  dets['anchor'][i, :n] = 100 * g + np.arange(n)
  dets['prob'][i, :n] = 0.5
blob = torch.from_numpy(shard.pack_blob(dets, counts))
gathered = shard.allgather_blob(blob, world)
gd, gc = shard.unpack_global(gathered.numpy(), GLOBAL_B, world, bmax, MAX_DETS)
np.savez(os.path.join(os.environ['OUT_DIR'], 'rank%d.npz' % rank), dets=gd, counts=gc)
dist.destroy_process_group()
