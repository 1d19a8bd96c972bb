#!/bin/bash
# ncu evidence: per-launch device times of whole steps (bench batch sizes), full captures of the three
# kernels that dominate a step.  Outputs under gpurun_out/.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_plain_nfs.csv python tools/prof_step.py plain_nfs 57 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_bmcnet_nfs.csv python tools/prof_step.py bmcnet_nfs 38 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:conv_slabt -s 8 -c 1 -f -o gpurun_out/slabt python tools/prof_step.py plain_nfs 57 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:bie_front -s 3 -c 1 -f -o gpurun_out/front python tools/prof_step.py plain_nfs 57 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:att_fold -s 3 -c 1 -f -o gpurun_out/fold python tools/prof_step.py plain_nfs 57 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 2 -c 1 -f -o gpurun_out/enc python -c "
import torch, sys
sys.path.insert(0, '.')
from bmcnet_esr_b200.dataloader import encodings as G
n = 100_000_000
xs = torch.rand(n, device='cuda') * 80; ys = torch.rand(n, device='cuda') * 45
ps = (torch.rand(n, device='cuda') < 0.5).float() * 2 - 1
for _ in range(4): G.events_to_channels(xs, ys, ps, sensor_size=(45, 80))
torch.cuda.synchronize()
" > /dev/null 2>&1; echo "rc=$?"
for wl in plain_nfs bmcnet_nfs; do B=57; [ $wl = bmcnet_nfs ] && B=38
BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py $wl $B 6 2>&1 | grep -E "optime" > gpurun_out/optimes_$wl.txt; tail -1 gpurun_out/optimes_$wl.txt; done
ls -la gpurun_out/*.ncu-rep
