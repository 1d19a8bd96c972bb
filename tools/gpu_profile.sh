#!/bin/bash
# the BMC_* switches below exist only in the measurement library (python -m bmcnet_esr_b200.build --measure)
export BMC_B200_LIB=${BMC_B200_LIB:-$PWD/bmcnet_esr_b200/libbmc_b200_measure.so}
# ncu evidence: per-launch device times of whole steps (bench batch sizes), full captures of the
# kernels that dominate a step, per-op CUDA-event timings.  Text summaries under gpurun_out/prof/.
mkdir -p gpurun_out/prof
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/prof/launches_plain_nfs_B95.csv python tools/prof_step.py plain_nfs 95 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/prof/launches_bmcnet_nfs_B76.csv python tools/prof_step.py bmcnet_nfs 76 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:conv_slabt -s 3 -c 1 -f -o gpurun_out/prof/slabt_mix python tools/prof_step.py plain_nfs 95 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:conv_slab2 -s 1 -c 1 -f -o gpurun_out/prof/slab2_plain3x3 python tools/time_conv.py 95 2 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:bie_front -s 3 -c 1 -f -o gpurun_out/prof/front python tools/prof_step.py plain_nfs 95 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:att_fold -s 3 -c 1 -f -o gpurun_out/prof/fold python tools/prof_step.py plain_nfs 95 2 > /dev/null 2>&1; echo "rc=$?"
cat > /tmp/enc_prof.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from bmcnet_esr_b200.dataloader import encodings as G
n = 100_000_000
xs = torch.rand(n, device='cuda') * 80; ys = torch.rand(n, device='cuda') * 45
ps = (torch.rand(n, device='cuda') < 0.5).float() * 2 - 1
ts = torch.sort(torch.rand(n, device='cuda'))[0]
for _ in range(3): G.events_to_channels(xs, ys, ps, sensor_size=(45, 80))
for _ in range(3): G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(45, 80))
torch.cuda.synchronize()
PY
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 2 -c 1 -f -o gpurun_out/prof/enc python /tmp/enc_prof.py > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 5 -c 1 -f -o gpurun_out/prof/vox python /tmp/enc_prof.py > /dev/null 2>&1; echo "rc=$?"
for f in slabt_mix slab2_plain3x3 front fold enc vox; do python tools/ncu_summary.py gpurun_out/prof/$f.ncu-rep > gpurun_out/prof/ncu_full_$f.txt 2>&1; done
for f in launches_plain_nfs_B95 launches_bmcnet_nfs_B76; do python tools/launch_summary.py gpurun_out/prof/$f.csv > gpurun_out/prof/$f.txt 2>&1; done
for wl in plain_nfs bmcnet_nfs; do B=95; [ $wl = bmcnet_nfs ] && B=76
BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py $wl $B 6 2>&1 | grep -E "optime" > gpurun_out/prof/optimes_${wl}_B$B.txt; tail -1 gpurun_out/prof/optimes_${wl}_B$B.txt; done
rm -f gpurun_out/prof/*.ncu-rep.tmp; ls -la gpurun_out/prof/
