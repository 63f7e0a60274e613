#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/desc_test tools/desc_test.cu > /dev/null 2>&1
timeout 120 /tmp/desc_test > gpurun_out/r2c_desc.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_desc.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_bias tools/mma_bias.cu > /dev/null 2>&1
timeout 120 /tmp/mma_bias > gpurun_out/r2c_mma_bias.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_mma_bias.log
: > gpurun_out/r2c_fire.log
export SQDET_FUSED_FIRE=2
for i in 0 1 2 3 4 5 6 7; do
  timeout 90 python tests/debug_fire.py $i >> gpurun_out/r2c_fire.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_fire.log
done
for i in 0 3 4 5 6 7; do
  timeout 120 python tests/debug_fire.py $i 8 24 78 >> gpurun_out/r2c_fire.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_fire.log
done
export SQDET_FUSED_FIRE=1
timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2c_perop_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_perop_fused.log
SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2c_perop_fused_dbg.log 2>&1
SQDET_FUSED_FIRE=2 timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2c_perop_fused_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_perop_fused_all.log
SQDET_FUSED_FIRE=2 SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2c_perop_fused_all_dbg.log 2>&1
cat gpurun_out/r2c_fire.log | grep -v "^rc=0"
tail -22 gpurun_out/r2c_perop_fused.log
