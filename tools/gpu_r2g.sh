#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/desc_test tools/desc_test.cu > /dev/null 2>&1
timeout 120 /tmp/desc_test 2>&1 | grep fence_cost > gpurun_out/r2g_fence.log
(timeout 600 python -m pytest tests/test_gpu_fire.py tests/test_gpu_kernels.py -q -x --timeout 200 --timeout-method thread) > gpurun_out/r2g_tests.log 2>&1
for cfg in "SQDET_POOL_FAST=1" "SQDET_FF_SQCAT=0" "SQDET_POOL_FAST=0"; do
  env $cfg timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2g_perop.tmp 2>&1
  echo "$cfg :: $(grep -E '^fire2 |^fire3 |^pool3|^pool5|^total' gpurun_out/r2g_perop.tmp | awk '{printf "%s %s  ", $1, $2}')" >> gpurun_out/r2g_sweep.log
done
SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 2>&1 | grep fire_tc | tail -2 > gpurun_out/r2g_dbg.log
cat gpurun_out/r2g_fence.log; tail -3 gpurun_out/r2g_tests.log; cat gpurun_out/r2g_sweep.log
