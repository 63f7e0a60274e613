"""Batch sharding across the GPUs of one box + the single all-gather of detections.

The reference is single-device (`tf.device('/gpu:{}')`, src/nets/squeezeDet.py:21).
Every image is independent end to end (conv stack per sample; filter_prediction per
image, src/demo.py:198, src/eval.py:81-87), so the path shards over the batch with NO
data-path exchange; the only collective is ONE all-gather of the fixed-size result
blob each rank's engine already lays out contiguously on the device:

    [B_local_max * max_dets * 28 B  sqdet_det records][B_local_max * 4 B  int32 counts]

Launch: one process per GPU (torchrun), NCCL over NVLink/NVSwitch on GPUs, gloo on
CPU for the host-logic tests.  The payload is a few KB per rank (latency-bound), so it
is issued on the compute stream right behind the filter kernel.
"""
from __future__ import annotations

import numpy as np

from ._lib import DET_DTYPE


def shard_sizes(batch, world):
  """Balanced contiguous split: the first batch % world ranks get one more image
  (b=20, N=8 -> 3,3,3,3,2,2,2,2)."""
  base, extra = divmod(int(batch), int(world))
  return [base + (1 if r < extra else 0) for r in range(world)]


def shard_ranges(batch, world):
  out, lo = [], 0
  for s in shard_sizes(batch, world):
    out.append((lo, lo + s))
    lo += s
  return out


def blob_nbytes(batch_local_max, max_dets):
  return batch_local_max * max_dets * DET_DTYPE.itemsize + batch_local_max * 4


def pack_blob(dets, counts):
  """Host-side equivalent of the engine's device layout (tests / CPU path)."""
  d = np.ascontiguousarray(dets).view(np.uint8).reshape(-1)
  c = np.ascontiguousarray(counts, dtype=np.int32).view(np.uint8).reshape(-1)
  return np.concatenate([d, c])


def unpack_blob(blob, batch_local_max, max_dets):
  blob = np.asarray(blob, dtype=np.uint8).reshape(-1)
  nd = batch_local_max * max_dets * DET_DTYPE.itemsize
  dets = blob[:nd].view(DET_DTYPE).reshape(batch_local_max, max_dets)
  counts = blob[nd:nd + 4 * batch_local_max].view(np.int32)
  return dets, counts


def allgather_blob(blob, world, group=None):
  """ONE collective: every rank contributes its blob (1-D uint8 torch tensor, device
  or CPU) and receives [world, nbytes]."""
  import torch
  import torch.distributed as dist
  out = torch.empty((world, blob.numel()), dtype=torch.uint8, device=blob.device)
  if world == 1:
    out[0].copy_(blob)
    return out
  dist.all_gather_into_tensor(out.view(-1), blob, group=group)
  return out


def unpack_global(gathered, global_batch, world, batch_local_max, max_dets):
  """[world, nbytes] uint8 -> (dets [global_batch, max_dets], counts [global_batch]),
  dropping each rank's padding images."""
  gathered = np.asarray(gathered)
  dets = np.zeros((global_batch, max_dets), DET_DTYPE)
  counts = np.zeros((global_batch,), np.int32)
  for r, (lo, hi) in enumerate(shard_ranges(global_batch, world)):
    d, c = unpack_blob(gathered[r], batch_local_max, max_dets)
    dets[lo:hi] = d[:hi - lo]
    counts[lo:hi] = c[:hi - lo]
  return dets, counts


class CudaView:
  """Zero-copy torch view of a raw device pointer (e.g. the engine's result blob)
  through __cuda_array_interface__."""

  def __init__(self, ptr, nbytes):
    self.__cuda_array_interface__ = {
        'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
        'version': 2, 'strides': None}


def device_blob_tensor(ptr, nbytes, device):
  import torch
  return torch.as_tensor(CudaView(ptr, nbytes), device=torch.device('cuda', device))
