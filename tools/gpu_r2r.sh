#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2r_tests.log 2>&1
timeout 300 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
tail -4 gpurun_out/r2r_tests.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2r_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 4), d['clocks'])
print([(r['op'], r['ms']) for r in d['per_op']])
PY
