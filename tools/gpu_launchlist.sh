#!/bin/bash
mkdir -p gpurun_out
for wl in ${WORKLOADS:-plain_nfs}; do
  timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_$wl.csv \
      python tools/prof_step.py $wl ${BATCH:-19} 3 > gpurun_out/prof_step_$wl.log 2>&1
  echo "launch list $wl rc=$?"; tail -2 gpurun_out/prof_step_$wl.log
done
timeout 600 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -5
