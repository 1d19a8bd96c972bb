#!/bin/bash
# ncu evidence: per-launch device times of whole steps, and one full capture of the dominant kernel.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for wl in plain_nfs bmcnet_nfs; do
  timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_$wl.csv \
      python tools/prof_step.py $wl 19 3 > gpurun_out/prof_step_$wl.log 2>&1
  echo "launch list $wl rc=$?"
done
timeout 600 $NCU --set full --import-source on -k regex:conv_slab -s 2 -c 1 -f -o gpurun_out/slab_4job \
    python tools/prof_conv.py 0 19 4 9 4 > gpurun_out/prof_conv.log 2>&1
echo "full capture rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:conv_slab -s 2 -c 1 -f -o gpurun_out/slab_2job \
    python tools/prof_conv.py 0 19 2 9 4 >> gpurun_out/prof_conv.log 2>&1
timeout 300 python tools/gpu_diag.py convperf > gpurun_out/convperf.log 2>&1
tail -30 gpurun_out/convperf.log
