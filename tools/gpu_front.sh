#!/bin/bash
# A/B loop for bie_front_tc: parity (model + kernel tests on the product library), per-op timings and the kernel's
# per-role cycle counters (measurement library), headline-only bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
export BMC_B200_LIB=$PWD/bmcnet_esr_b200/libbmc_b200_measure.so
BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py plain_nfs 95 6 2>&1 | grep -E "optime|Error|error" > gpurun_out/optimes_plain_nfs.txt
grep -E "bie_front|total" gpurun_out/optimes_plain_nfs.txt | tail -3
BMC_FRONT_PROF=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py plain_nfs 95 4 2>&1 | grep frontprof | head -2
unset BMC_B200_LIB
timeout 600 python bench.py --steps 20 --warmup 5 --only-headline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms', d['ms_per_step'], 'sustained', d.get('value_sustained'), 'e2e', d['e2e']['value'])"
