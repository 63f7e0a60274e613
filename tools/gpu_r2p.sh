#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2p_bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 4), d['clocks'])
PY
cat gpurun_out/r2p_smoke.log | tail -2
