#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_fire.py -q -x --timeout 200 --timeout-method thread) > gpurun_out/r2q_tests.log 2>&1
for cfg in "SQDET_FF_SQSPLIT=1" "SQDET_FF_SQSPLIT=0" "SQDET_FF_SQSPLIT=1 SQDET_FF_SQCAT=1"; do
  env $cfg timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2q_perop.tmp 2>&1
  echo "$cfg :: $(grep -E '^fire2 |^fire3 |^total' gpurun_out/r2q_perop.tmp | awk '{printf "%s %s  ", $1, $2}')" >> gpurun_out/r2q_sweep.log
done
SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 2>&1 | grep fire_tc | tail -2 > gpurun_out/r2q_dbg.log
tail -2 gpurun_out/r2q_tests.log; cat gpurun_out/r2q_sweep.log; cat gpurun_out/r2q_dbg.log
