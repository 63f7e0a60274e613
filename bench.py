#!/usr/bin/env python
"""bench.py — images/sec of the SqueezeDet inference hot path on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the CPU restatement of the reference's path

One "step" = one pass of the hot path (backbone + ConvDet + interpret_output +
filter_prediction/NMS) over one batch of synthetic 1242x375 images, b=20 PER GPU (weak
scaling: the batch shards over GPUs with no data-path exchange; one all-gather of the
filtered detections per step when N > 1).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'images/sec (1242x375, b=20)'
NETS = {'squeezeDet': ('SqueezeDet', 'kitti_squeezeDet_config'),
        'squeezeDet+': ('SqueezeDetPlus', 'kitti_squeezeDetPlus_config'),
        'vgg16': ('VGG16ConvDet', 'kitti_vgg16_config'),
        'resnet50': ('ResNet50ConvDet', 'kitti_res50_config')}
GRIDS = {'squeezeDet': (24, 78), 'squeezeDet+': (22, 76), 'vgg16': (24, 78), 'resnet50': (24, 78)}


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--net', default='squeezeDet', choices=sorted(NETS))
  ap.add_argument('--batch', type=int, default=20, help='images per GPU per step')
  ap.add_argument('--width', type=int, default=1242)
  ap.add_argument('--height', type=int, default=375)
  ap.add_argument('--math', default='tc', choices=['tc', 'simt'])
  ap.add_argument('--cpu-sample', type=int, default=20,
                  help='images in the cpu_baseline sample (N=1, rank 0)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  return ap.parse_args()


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as f:
      p = json.load(f)
    return dict(hbm_gbs=float(p['hbm_gbs']), tflops=float(p['bf16_tflops']),
                tflops_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                source='measured')
  except Exception:
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source='fallback')


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
  Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.rows = []
    self.proc = None
    self.thread = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
           '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
      self.proc = None
      return
    self.thread = threading.Thread(target=self._read, daemon=True)
    self.thread.start()

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append(line.strip())

  def stop(self):
    if not self.proc:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.12)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      parts = [p.strip() for p in r.split(',')]
      if len(parts) < 6:
        continue
      try:
        sm.append(float(parts[0]))
        smax.append(float(parts[1]))
      except ValueError:
        continue
      for nm, v in zip(names, parts[2:6]):
        if v.lower().startswith('active'):
          reasons.add(nm)
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': max(smax) if smax else None, 'reasons': sorted(reasons),
            'samples': len(sm)}


def host_threads():
  """Host threads this process may actually use (cgroup/affinity aware; os.cpu_count()
  over-reports on shared GPU hosts and oversubscribed torch is 50x slower)."""
  try:
    n = len(os.sched_getaffinity(0))
  except AttributeError:
    n = os.cpu_count() or 1
  try:   # cgroup v2 cpu.max quota
    with open('/sys/fs/cgroup/cpu.max') as f:
      quota, period = f.read().split()
    if quota != 'max':
      n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
  except Exception:
    pass
  return n


def build_mc(args):
  from squeezedet_b200 import config as cfg
  mc = getattr(cfg, NETS[args.net][1])()
  mc.IMAGE_WIDTH, mc.IMAGE_HEIGHT, mc.BATCH_SIZE = args.width, args.height, args.batch
  mc.ANCHOR_BOX = cfg.set_anchors(mc)
  mc.ANCHORS = len(mc.ANCHOR_BOX)
  return mc


# ------------------------------------------------------------------------------------------
def cpu_port_rate(args, n_images, repeats=1, threads=None):
  """The oracle's torch-CPU restatement + the oracle's numpy interpret/filter, timed on the
  host cores.  (The ONLY place bench.py touches oracle/: the checker used as a baseline.)"""
  import oracle
  from oracle.torch_port import TorchForward
  from squeezedet_b200.utils import synth
  mc = build_mc(args)
  weights = synth.synthetic_weights(oracle.param_specs(args.net), seed=0)
  if threads is None:
    # all the host threads it can use -- but measured, not assumed: on shared GPU hosts the
    # visible CPU count exceeds the usable one and oversubscribed torch collapses.
    import torch
    probe = synth.synthetic_images(1, args.height, args.width, seed=1)
    best, threads = None, 1
    nmax = host_threads()
    for cand in sorted({nmax, min(nmax, 64), min(nmax, 32), min(nmax, 16), min(nmax, 8)},
                       reverse=True):
      f = TorchForward(args.net, weights, threads=cand)
      f(probe)
      t0 = time.perf_counter()
      f(probe)
      dt = time.perf_counter() - t0
      if best is None or dt < best:
        best, threads = dt, cand
  fwd = TorchForward(args.net, weights, threads=threads)
  images = synth.synthetic_images(n_images, args.height, args.width, seed=1234)

  def one_pass():
    chunk = 4
    for i in range(0, n_images, chunk):
      preds = fwd(images[i:i + chunk])
      boxes, probs, cls = oracle.interpret_output(
          preds, mc.ANCHOR_BOX, mc.CLASSES, mc.ANCHOR_PER_GRID, mc.IMAGE_WIDTH,
          mc.IMAGE_HEIGHT, mc.EXP_THRESH)
      for j in range(len(probs)):
        oracle.filter_prediction(boxes[j], probs[j], cls[j], mc.CLASSES, mc.TOP_N_DETECTION,
                                 mc.PROB_THRESH, mc.NMS_THRESH)
  return one_pass, threads


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return            # other ranks exit 0 without work
  sample = 4          # images per step: bounded so K steps end within a few minutes
  one_pass, threads = cpu_port_rate(args, sample)
  for _ in range(max(args.warmup, 1)):
    one_pass()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    one_pass()
  dt = time.perf_counter() - t0
  value = sample * args.steps / dt
  what = ('%d synthetic %dx%d images/step through the oracle restatement (torch-CPU convs '
          '+ numpy interpret_output/filter_prediction); TF-1.0 itself is not installable'
          % (sample, args.width, args.height))
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/sec',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(args), 'net': args.net, 'batch_per_step': sample},
      'cpu_baseline': {'value': value, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
                       'sample': what},
      'e2e': {'value': value, 'unit': 'images/sec', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'gpu_launches': 0}))


def workload_name(args):
  return '%s inference, synthetic %dx%d, batch %d per GPU, random (calibrated) weights' % (
      args.net, args.width, args.height, args.batch)


# ------------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist
  from squeezedet_b200 import _lib, nets, shard
  from squeezedet_b200.utils import synth

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus and world > 1:
    raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
  if _lib.device_count() < 1:
    raise SystemExit('bench.py: no CUDA device visible; the engine has no CPU fallback')
  torch.cuda.set_device(local)
  if world > 1:
    # stdout carries exactly one JSON line: keep NCCL's version banner off it
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
      os.environ['NCCL_DEBUG'] = 'WARN'
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', local))

  mc = build_mc(args)
  math_mode = _lib.MATH_TF32X3_TC if args.math == 'tc' else _lib.MATH_FP32_SIMT
  model = getattr(nets, NETS[args.net][0])(mc, local, math_mode=math_mode)
  model.load_weights(synth.synthetic_weights(synth.model_param_specs(model), seed=0))
  B = args.batch

  # host inputs in pinned memory (e2e path) and a device-resident copy (`value` path)
  pinned = _lib.PinnedArray((B, args.height, args.width, 3), np.float32)
  pinned.array[...] = synth.synthetic_images(B, args.height, args.width, seed=1234 + rank)
  x_dev = torch.from_numpy(pinned.array).cuda(local)
  stream = torch.cuda.Stream(device=local)
  sptr = stream.cuda_stream
  res = model.results_device()
  nbytes_blob = shard.blob_nbytes(B, res['max_dets'])
  blob = shard.device_blob_tensor(res['dets'], nbytes_blob, local)
  gathered = torch.empty((world, nbytes_blob), dtype=torch.uint8, device=x_dev.device)

  def step_device():
    model.forward_device(x_dev.data_ptr(), sptr)
    if world > 1:
      with torch.cuda.stream(stream):
        dist.all_gather_into_tensor(gathered.view(-1), blob)

  # end-to-end through the C ABI with HOST buffers, pipelined two deep (sqdet_submit /
  # sqdet_wait): every step copies ITS inputs H2D from pinned memory and ITS records D2H.
  dets_host = [_lib.PinnedArray((B, res['max_dets']), _lib.DET_DTYPE) for _ in range(2)]
  counts_host = [_lib.PinnedArray((B,), np.int32) for _ in range(2)]
  pinned_u8 = _lib.PinnedArray((B, args.height, args.width, 3), np.uint8)
  pinned_u8.array[...] = np.random.default_rng(99 + rank).integers(
      0, 256, pinned_u8.array.shape, dtype=np.uint8)
  lib = _lib.load()
  e2e_state = {'i': 0, 'kind': _lib.IMG_U8, 'src': pinned_u8.ptr}

  def step_e2e():
    i = e2e_state['i']
    _lib.check(lib.sqdet_submit(model._engine, e2e_state['src'], e2e_state['kind'],
                                dets_host[i & 1].ptr, counts_host[i & 1].ptr))
    if i >= 1:
      _lib.check(lib.sqdet_wait(model._engine))        # results of step i-1 are on the host
      if world > 1:
        with torch.cuda.stream(stream):
          dist.all_gather_into_tensor(gathered.view(-1), blob)
    e2e_state['i'] = i + 1

  def drain_e2e():
    while True:
      try:
        _lib.check(lib.sqdet_wait(model._engine))
      except _lib.SqdetError:
        break
    e2e_state['i'] = 0
    torch.cuda.synchronize()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
      fn()
    e1.record(stream)
    stream.synchronize()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=x_dev.device)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  W = max(args.warmup, 3)
  for _ in range(W):
    step_device()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  ms_total = timed(step_device, args.steps)
  ms_per_step = ms_total / args.steps
  value = world * B / (ms_per_step * 1e-3)

  def timed_e2e(kind, src):
    e2e_state.update(kind=kind, src=src)
    for _ in range(3):
      step_e2e()
    drain_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
      step_e2e()
    drain_e2e()                                   # last step's results delivered
    dt = time.perf_counter() - t0
    barrier()
    ms = torch.tensor([dt * 1e3], device=x_dev.device)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / args.steps

  ms_e2e = timed_e2e(_lib.IMG_U8, pinned_u8.ptr)
  e2e_value = world * B / (ms_e2e * 1e-3)
  ms_e2e_f32 = timed_e2e(_lib.IMG_F32, pinned.ptr)
  # extra loaded steps so the 100 ms nvidia-smi sampler sees the GPU under this workload
  if rank == 0:
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
      model.forward_device(x_dev.data_ptr(), sptr)     # no collective: rank-local
      stream.synchronize()
  clocks = sampler.stop() if rank == 0 else None

  # ---- roofline of the dominant kernel: per-op CUDA-event times, measured live ----------
  roofline = None
  per_op = None
  if rank == 0:
    peaks = measured_peaks()
    acc = None
    reps = 5
    for _ in range(reps):
      rows = model.forward_profiled(x_dev.data_ptr(), sptr)
      t = np.array([ms for _, ms in rows])
      acc = t if acc is None else acc + t
    acc /= reps
    table = model.op_table()
    per_op = [{'op': nm, 'ms': round(float(ms), 4), 'gflop': round(fl / 1e9, 3),
               'mbytes': round(by / 1e6, 2)} for (nm, fl, pa, by), ms in zip(table, acc)]
    top = int(np.argmax(acc))
    nm, fl, pa, by = table[top]
    ridge = peaks['tflops_sustained'] * 1e12 / (peaks['hbm_gbs'] * 1e9)
    sec = float(acc[top]) * 1e-3
    if fl / max(by, 1) >= ridge:
      ach, peak, unit, bound = fl / sec / 1e12, peaks['tflops_sustained'], 'TFLOP/s', 'tensor'
    else:
      ach, peak, unit, bound = by / sec / 1e9, peaks['hbm_gbs'], 'GB/s', 'hbm'
    tot_by = sum(r[3] for r in table)
    tot_fl = sum(r[1] for r in table)
    traffic = None
    try:
      with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
        traffic = json.load(f)['bytes_per_op'].get(nm)
    except Exception:
      pass
    roofline = {'kernel': nm, 'bound': bound, 'achieved': ach, 'peak': peak, 'unit': unit,
                'frac': ach / peak, 'traffic': traffic,
                'traffic_source': 'profiles/r1_traffic.json (ncu dram bytes, same workload)' if traffic else None, 'peak_source': peaks['source'],
                'kernel_ms': float(acc[top]), 'kernel_share_of_step': float(acc[top] / acc.sum()),
                # context for FFMA kernels (the fused first layer): algorithmic TFLOP/s next to the
                # fp32 SIMT peak of this part (SMs x 128 lanes x 2 flop x SM clock)
                'achieved_tflops': fl / sec / 1e12,
                'fp32_simt_peak_tflops': 148 * 128 * 2 * (clocks['sm_mhz'] if clocks and clocks.get('sm_mhz') else 1965.0) * 1e6 / 1e12,
                'algorithmic_bytes_per_launch': by, 'algorithmic_flops_per_launch': fl,
                'whole_step': {'algorithmic_gbytes': tot_by / 1e9, 'gflop': tot_fl / 1e9,
                               'hbm_gbs': tot_by / (ms_per_step * 1e-3) / 1e9,
                               'hbm_frac': tot_by / (ms_per_step * 1e-3) / 1e9 / peaks['hbm_gbs'],
                               'tflops': tot_fl / (ms_per_step * 1e-3) / 1e12}}

  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    n = args.cpu_sample
    one_pass, threads = cpu_port_rate(args, n)
    one_pass()                                   # warm-up
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
      one_pass()
      reps += 1
    dt = time.perf_counter() - t0
    cpu_baseline = {'value': n * reps / dt, 'unit': 'images/sec', 'cores': threads,
                    'kind': 'port',
                    'sample': '%d passes over %d synthetic %dx%d images: oracle restatement '
                              '(torch-CPU convs, all host threads) + numpy interpret/filter'
                              % (reps, n, args.width, args.height)}

  if rank == 0:
    h2d = int(pinned_u8.array.nbytes)
    d2h = int(dets_host[0].array.nbytes + counts_host[0].array.nbytes)
    launches = model.launches_per_forward()
    out = {
        'metric': METRIC, 'value': value, 'unit': 'images/sec', 'n_gpus': world,
        'steps': args.steps, 'warmup': W, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 (3xTF32 tcgen05, fp32 accumulate)' if args.math == 'tc' else 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_name(args), 'net': args.net,
                   'global_batch': world * B, 'image': [args.height, args.width],
                   'parallelism': 'batch-sharded x%d, one all-gather of detections' % world,
                   'l2': 'no flush: %.1f GB of activations stream through the 126 MB L2 '
                         'every step' % (sum(r[3] for r in model.op_table()) / 1e9),
                   'math': args.math},
        'e2e': {'value': e2e_value, 'unit': 'images/sec', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e,
                'input': 'uint8 BGR images in pinned host memory; `- mc.BGR_MEANS` '
                         '(demo.py:190) runs on the GPU; sqdet_submit/sqdet_wait, 2 in flight',
                'timer': 'host wall clock around K submits + final wait (covers H2D, kernels, '
                         'D2H), max over ranks'},
        'e2e_f32_feed': {'value': world * B / (ms_e2e_f32 * 1e-3), 'unit': 'images/sec',
                         'h2d_bytes_per_step': int(pinned.array.nbytes),
                         'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e_f32,
                         'input': 'float32 mean-subtracted images (the reference feed_dict '
                                  'payload), PCIe-bound'},
        'gpu_launches': launches * args.steps,
        'launches_per_step': launches,
        'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'per_op': per_op,
    }
    print(json.dumps(out))
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
