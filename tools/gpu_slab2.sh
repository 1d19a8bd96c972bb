#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
# A/B of the slab convolution kernels: parity tests, per-launch timing, bench lines, per-op timings.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_metrics.py -x -q > gpurun_out/pytest_slab2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_slab2.log
tail -15 gpurun_out/pytest_slab2.log
for v in 1 0; do for g in 148 16; do BMC_CONV_SLAB2=$v BMC_SLABT_GRID=$g timeout 120 python tools/time_conv.py 57 2 2>&1 | grep -E "conv3x3|rror|timeout" | sed "s/^/slab2=$v /" | tee -a gpurun_out/convgrid.txt; done; done
for wl in ${WORKLOADS:-plain_nfs bmcnet_nfs}; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --cpu-steps 2 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  tail -c 400 gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$wl.json'))
    print('$wl value %.0f e2e %.0f ms/step %.3f roofline %.0f TF (%.2f)' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac']))
except Exception as e: print('no json', e)
PY
done
for wl in plain_nfs bmcnet_nfs; do B=57; [ $wl = bmcnet_nfs ] && B=38
BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py $wl $B 6 2>&1 | grep -E "optime" > gpurun_out/optimes_${wl}.txt; tail -1 gpurun_out/optimes_${wl}.txt; done
