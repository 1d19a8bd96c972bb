#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
BMC_CONV_SLABT=0 timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -1
BMC_CONV_SLABT=0 bash tools/gpu_quick.sh 2>&1 | grep -E "value"
