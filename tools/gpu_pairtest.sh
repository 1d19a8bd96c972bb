#!/bin/bash
mkdir -p gpurun_out
BMC_CONV_PAIR=0 timeout 240 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | grep -E "assert|Error|passed|failed" | head -6
BMC_CONV_PAIR=0 BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py plain_nfs 19 6 2>&1 | grep -E "optime" > gpurun_out/optimes_plain_nfs.txt
head -6 gpurun_out/optimes_plain_nfs.txt; tail -1 gpurun_out/optimes_plain_nfs.txt
BMC_CONV_PAIR=0 bash tools/gpu_quick.sh 2>&1 | grep -E "value"
