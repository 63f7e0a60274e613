#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fire.py tests/test_gpu_adversarial.py -q -x --timeout 200 --timeout-method thread) > gpurun_out/r2i_tests.log 2>&1
for cfg in "SQDET_TC_2SPLIT=2" "SQDET_TC_2SPLIT=0" "SQDET_TC_2SPLIT=1"; do
  env $cfg timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2i_perop.tmp 2>&1
  echo "$cfg :: $(grep -E '^fire|^conv12|^total' gpurun_out/r2i_perop.tmp | awk '{printf "%s %s  ", $1, $2}')" >> gpurun_out/r2i_sweep.log
done
for net in vgg16 resnet50 squeezeDet+; do
  b=8; [ "$net" = "squeezeDet+" ] && b=20
  for m in 2 0 1; do
    SQDET_TC_2SPLIT=$m timeout 200 python tests/debug_forward.py $net $b > gpurun_out/r2i_tmp.log 2>&1
    echo "$net 2split=$m $(grep '^total' gpurun_out/r2i_tmp.log)" >> gpurun_out/r2i_sweep.log
  done
done
tail -3 gpurun_out/r2i_tests.log; cat gpurun_out/r2i_sweep.log
