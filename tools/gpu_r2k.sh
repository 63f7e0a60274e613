#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 --timeout-method thread) > gpurun_out/r2k_multi_tests.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err
tail -4 gpurun_out/r2k_multi_tests.log
python - <<'PY'
import json
try:
  d = json.load(open('gpurun_out/r2k_bench_2gpu.json'))
  print('2gpu value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'gather', d.get('gather_verified'))
  s = d.get('strong_scaling'); print('strong', s and (round(s['value']), s['ms_per_step'], s['shards']))
  for o in d['other_configs']: print(o['net'], o['scaling'], round(o['value']), round(o['ms_per_step'], 3))
except Exception as e:
  print('unreadable', e); print(open('gpurun_out/r2k_bench_2gpu.err').read()[-1500:])
PY
