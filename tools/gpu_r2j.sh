#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -x --timeout 300 --timeout-method thread) > gpurun_out/r2j_tests.log 2>&1
for cfg in "SQDET_TC_SPLITK_ROWS=0" "SQDET_TC_SPLITK_ROWS=1" "SQDET_TC_SEG=4" "SQDET_TC_SEG=6"; do
  env $cfg timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2j_perop.tmp 2>&1
  echo "$cfg :: $(grep -E '^fire4|^fire1|^conv12|^total' gpurun_out/r2j_perop.tmp | awk '{printf "%s %s  ", $1, $2}')" >> gpurun_out/r2j_sweep.log
done
tail -3 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_sweep.log
