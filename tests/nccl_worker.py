"""Worker for tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run).
Global batch G shards over the ranks (shard.py); every rank runs its shard through its own
engine with the all-gather captured IN the forward graph (sqdet_set_gather_in_forward), then
rank 0 also runs the whole global batch on one engine; the gathered records must be byte-equal."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

import oracle
from squeezedet_b200 import _lib, nets, shard
from squeezedet_b200 import config as cfg
from squeezedet_b200.utils import synth

import traceback


def _excepthook(tp, val, tb):
  # torchrun's summary hides the workers' own tracebacks: leave them where the test can read them
  try:
    with open(os.path.join(os.environ.get('OUT_DIR', '.'), 'err_rank%s.txt' % os.environ.get('RANK', '?')), 'w') as f:
      traceback.print_exception(tp, val, tb, file=f)
  finally:
    traceback.print_exception(tp, val, tb)


sys.excepthook = _excepthook
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
G = int(os.environ.get('GLOBAL_BATCH', '5'))
W, H = 416, 128
torch.cuda.set_device(local)
# host channel for the engine communicator's 128-byte id and the final verdict: torch's own NCCL group
# (gloo resolves the container hostname, which GPU boxes do not always resolve)
dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))


def make_model(batch):
  mc = cfg.kitti_squeezeDet_config()
  mc.IMAGE_WIDTH, mc.IMAGE_HEIGHT, mc.BATCH_SIZE = W, H, batch
  grid = oracle.layer_table('squeezeDet', H, W)[-1][2]      # ConvDet grid of this image size
  mc.GRID_H, mc.GRID_W = grid[0], grid[1]
  mc.ANCHOR_BOX = cfg.set_anchors(mc)
  mc.ANCHORS = len(mc.ANCHOR_BOX)
  m = nets.SqueezeDet(mc, local)
  m.load_weights(synth.synthetic_weights(synth.model_param_specs(m), seed=0))
  return m


images = synth.synthetic_images(G, H, W, seed=77)
bmax = max(shard.shard_sizes(G, world))
lo, hi = shard.shard_ranges(G, world)[rank]
mine = np.zeros((bmax, H, W, 3), np.float32)
mine[:hi - lo] = images[lo:hi]                    # short shards are padded with zero images
model = make_model(bmax)
ident = [_lib.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ident, src=0)
model.comm_init(world, rank, ident[0], in_forward=True)
for _ in range(3):                                # graph capture, then replays
  dets, counts = model.detect_records(mine)
gathered = model.read_gathered()
gd, gc = shard.unpack_global(gathered, G, world, bmax, model.max_dets)
ok = True
if rank == 0:
  ref = make_model(G)
  rd, rc = ref.detect_records(images)
  ok = bool(np.array_equal(gc, rc) and np.array_equal(gd, rd))
  print('nccl_worker: world %d global batch %d byte-equal %s kept %d'
        % (world, G, ok, int(rc.sum())), flush=True)
np.savez(os.path.join(os.environ['OUT_DIR'], 'rank%d.npz' % rank), dets=gd, counts=gc)
flag = torch.tensor([1 if ok else 0], device=torch.device('cuda', local))
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
model.comm_destroy()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 3)
