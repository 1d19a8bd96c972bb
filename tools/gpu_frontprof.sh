#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
mkdir -p gpurun_out
BMC_FRONT_PROF=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py plain_nfs 19 4 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
WORKLOADS=plain_nfs bash tools/gpu_launchlist.sh | head -3
