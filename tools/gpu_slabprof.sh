#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
mkdir -p gpurun_out
BMC_SLAB_PROF=1 timeout 300 python tools/prof_conv.py 0 19 2 9 6 2>&1 | grep -E "slabprof|done|rror" 
timeout 600 python tools/gpu_diag.py convperf 2>&1 | grep -E "taps=9.*B=19|taps=9.*B=16"
bash tools/gpu_optimes.sh
