#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_fire.py -q -x --timeout 200 --timeout-method thread) > gpurun_out/r2l_tests.log 2>&1
for perm in 3210 1032 3102 0132 2301; do
  SQDET_FF_PERM=$perm timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2l_perop.tmp 2>&1
  echo "SQDET_FF_PERM=$perm :: $(grep -E '^fire2 |^fire3 |^total' gpurun_out/r2l_perop.tmp | awk '{printf "%s %s  ", $1, $2}')" >> gpurun_out/r2l_sweep.log
done
SQDET_FF_PERM=1032 timeout 200 python -m pytest tests/test_gpu_fire.py -q -x --timeout 200 --timeout-method thread > gpurun_out/r2l_tests_perm.log 2>&1
tail -2 gpurun_out/r2l_tests.log; tail -2 gpurun_out/r2l_tests_perm.log; cat gpurun_out/r2l_sweep.log
