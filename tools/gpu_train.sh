#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_dp.py -x -q > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_train.log
tail -5 gpurun_out/pytest_train.log
timeout 600 python tools/train_time.py ${BATCHES:-2 8} > gpurun_out/train_time.log 2>&1; grep -v "^  " gpurun_out/train_time.log | tail -12
