#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
for wl in ${WORKLOADS:-plain_nfs bmcnet_nfs}; do
BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py $wl ${BATCH:-19} 6 2>&1 | grep -E "optime|Error|error" > gpurun_out/optimes_$wl.txt
tail -1 gpurun_out/optimes_$wl.txt
done
bash tools/gpu_quick.sh 2>&1 | grep -E "value|rc=1"
