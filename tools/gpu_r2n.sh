#!/bin/bash
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2n_tests.log 2>&1
timeout 150 python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2n_perop.log 2>&1
SQDET_TC_DEBUG=1 timeout 150 python tests/debug_forward.py squeezeDet 20 2>&1 | grep fire_tc | tail -2 > gpurun_out/r2n_dbg.log
timeout 600 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/r2n_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 25 -c 25 -f -o /tmp/r2_full python tests/debug_forward.py squeezeDet 20 > gpurun_out/r2n_ncu_full.log 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw.csv 2>> gpurun_out/r2n_ncu_full.log
ls -la /tmp/r2_full.ncu-rep >> gpurun_out/r2n_ncu_full.log
tail -4 gpurun_out/r2n_tests.log
tail -20 gpurun_out/r2n_perop.log
