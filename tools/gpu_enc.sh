#!/bin/bash
# encoder parity tests + kernel timings.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoders.py -x -q > gpurun_out/pytest_enc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_enc.log
tail -5 gpurun_out/pytest_enc.log
for n in 1e8 4e8; do timeout 300 python tools/bench_enc.py $n 2>&1 | grep -E "^enc" | tee -a gpurun_out/bench_enc.txt; done
