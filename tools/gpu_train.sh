#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -k "graphed" > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_train.log
tail -30 gpurun_out/pytest_train.log
timeout 600 python tools/train_time.py ${BATCHES:-2 8} 2>&1 | tail -20
