#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
mkdir -p gpurun_out
for d in 0 1 2 3; do for g in 148 16; do BMC_SLAB2_DBG=$d BMC_CONV_SLAB2=1 BMC_SLABT_GRID=$g timeout 120 python tools/time_conv.py 57 2 2>&1 | grep -E "conv3x3|rror|slab2prof" | sed "s/^/dbg=$d /" | tee -a gpurun_out/slab2prof.txt; done; done
