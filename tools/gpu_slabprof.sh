#!/bin/bash
mkdir -p gpurun_out
BMC_SLAB_PROF=1 timeout 300 python tools/prof_conv.py 0 19 2 9 6 2>&1 | grep -E "slabprof|done|rror" 
timeout 600 python tools/gpu_diag.py convperf 2>&1 | grep -E "taps=9.*B=19|taps=9.*B=16"
bash tools/gpu_optimes.sh
