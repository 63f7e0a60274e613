#!/bin/bash
# 2-GPU call: multi-GPU tests (+ the worker run directly, its stderr kept)
mkdir -p gpurun_out /tmp/ncclw
OUT_DIR=/tmp/ncclw GLOBAL_BATCH=5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_worker.py > gpurun_out/r2o_worker.log 2>&1
cat /tmp/ncclw/err_rank*.txt >> gpurun_out/r2o_worker.log 2>/dev/null
(timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 --timeout-method thread) > gpurun_out/r2o_multi_tests.log 2>&1
tail -4 gpurun_out/r2o_multi_tests.log
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/r2o_worker.log | tail -30
