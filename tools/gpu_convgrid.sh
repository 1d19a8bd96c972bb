#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
mkdir -p gpurun_out
for g in 148 111 74 37 16; do BMC_SLABT_GRID=$g timeout 120 python tools/time_conv.py 57 2 2>&1 | grep conv3x3 | tee -a gpurun_out/convgrid.txt; done
