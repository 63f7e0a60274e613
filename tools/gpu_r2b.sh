#!/bin/bash
# one gpurun call: descriptor unit test, fused fire per shape, per-op timings fused / unfused
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/desc_test tools/desc_test.cu > /dev/null 2>&1
timeout 60 /tmp/desc_test > gpurun_out/r2b_desc.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_desc.log
: > gpurun_out/r2b_fire.log
export SQDET_FUSED_FIRE=2
for i in 0 1 2 3 4 5 6 7; do
  timeout 90 python tests/debug_fire.py $i >> gpurun_out/r2b_fire.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_fire.log
done
timeout 90 python tests/debug_fire.py 0 1 24 78 >> gpurun_out/r2b_fire.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_fire.log
timeout 90 python tests/debug_fire.py 3 2 47 156 >> gpurun_out/r2b_fire.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_fire.log
timeout 90 python tests/debug_fire.py 1 1 8 16 >> gpurun_out/r2b_fire.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_fire.log
export SQDET_FUSED_FIRE=1
SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2b_perop_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_perop_fused.log
SQDET_FUSED_FIRE=0 timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2b_perop_unfused.log 2>&1
SQDET_FUSED_FIRE=2 timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2b_perop_fused_all.log 2>&1
cat gpurun_out/r2b_desc.log gpurun_out/r2b_fire.log
grep -v "^\[" gpurun_out/r2b_perop_fused.log | tail -30
