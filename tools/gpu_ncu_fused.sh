#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none --cache-control none"
timeout 600 $NCU --set full --import-source on -k regex:att_fold -s 6 -c 1 -f -o gpurun_out/fold python tools/prof_step.py plain_nfs 19 4 > gpurun_out/ncu_fold.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:bie_front -s 6 -c 1 -f -o gpurun_out/front python tools/prof_step.py plain_nfs 19 4 > gpurun_out/ncu_front.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_warm_plain_nfs.csv python tools/prof_step.py plain_nfs 19 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
