#!/bin/bash
mkdir -p gpurun_out
perop() { grep -E "^fire[2-5] |^total" "$1" | awk '{printf "%s %s  ", $1, $2}'; echo; }
for cfg in "SQDET_FUSED_FIRE=1" "SQDET_FUSED_FIRE=1 SQDET_FF_NSQ=3" "SQDET_FUSED_FIRE=3" "SQDET_FUSED_FIRE=3 SQDET_FF_NQ=1" "SQDET_FUSED_FIRE=3 SQDET_FF_NSQ=2" "SQDET_FUSED_FIRE=0"; do
  env $cfg timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2d_perop.tmp 2>&1
  echo "$cfg :: $(perop gpurun_out/r2d_perop.tmp)" >> gpurun_out/r2d_sweep.log
done
SQDET_FUSED_FIRE=3 SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 2>&1 | grep fire_tc | tail -4 > gpurun_out/r2d_dbg.log
for pdl in 1 0; do
  SQDET_PDL=$pdl timeout 300 python bench.py --no-other-configs --no-cpu-baseline > gpurun_out/r2d_bench_pdl$pdl.json 2> gpurun_out/r2d_bench_pdl$pdl.err
done
(time timeout 1000 python -m pytest tests -m gpu -x -q) > gpurun_out/r2d_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cat gpurun_out/r2d_sweep.log
tail -5 gpurun_out/r2d_tests.log
python - <<'PY'
import json
for f in ('r2d_bench_pdl1', 'r2d_bench_pdl0', 'r2d_bench'):
  try:
    d = json.load(open('gpurun_out/%s.json' % f))
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'launches', d.get('launches_per_step'))
  except Exception as e:
    print(f, 'unreadable', e)
PY
