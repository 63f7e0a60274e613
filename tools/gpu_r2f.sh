#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_halo.py -q --timeout 120 --timeout-method thread) > gpurun_out/r2f_halo_tests.log 2>&1
for cfg in "SQDET_HALO_CONV=1" "SQDET_HALO_CONV=0" "SQDET_HALO_CONV=1 SQDET_HALO_KSPLIT=3" "SQDET_HALO_CONV=1 SQDET_HALO_KSPLIT=6" "SQDET_HALO_CONV=1 SQDET_HALO_NQ=2"; do
  env $cfg timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2f_perop.tmp 2>&1
  echo "$cfg :: $(grep -E '^conv12|^total' gpurun_out/r2f_perop.tmp | awk '{printf "%s %s  ", $1, $2}')" >> gpurun_out/r2f_sweep.log
done
SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 2>&1 | grep halo_tc | tail -2 > gpurun_out/r2f_dbg.log
for net in vgg16 resnet50; do
  for m in 1 2; do
    SQDET_HALO_CONV=$m timeout 200 python tests/debug_forward.py $net 8 > gpurun_out/r2f_${net}_halo$m.log 2>&1
    echo "$net halo=$m $(grep '^total' gpurun_out/r2f_${net}_halo$m.log)" >> gpurun_out/r2f_sweep.log
  done
done
tail -5 gpurun_out/r2f_halo_tests.log; cat gpurun_out/r2f_sweep.log; cat gpurun_out/r2f_dbg.log
