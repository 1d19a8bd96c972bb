#!/bin/bash
# Round-2 evidence: launch lists (bench batches and batch 1), --set full captures of the kernels that dominate a step and of
# the new training / encoder kernels, per-op CUDA-event timings, compute-sanitizer over the new kernels.
# Text summaries under gpurun_out/prof/ (copied to profiles/r02_* by hand).
mkdir -p gpurun_out/prof
NCU="ncu --clock-control none"
P=gpurun_out/prof
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_plain_nfs_B95.csv python tools/prof_step.py plain_nfs 95 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_bmcnet_nfs_B76.csv python tools/prof_step.py bmcnet_nfs 76 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_plain_nfs_B1.csv python tools/prof_step.py plain_nfs 1 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_bmcnet_nfs_B1.csv python tools/prof_step.py bmcnet_nfs 1 3 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:conv_slabt -s 3 -c 1 -f -o $P/slabt_mix python tools/prof_step.py plain_nfs 95 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:conv_slab2 -s 1 -c 1 -f -o $P/slab2_plain3x3 python tools/time_conv.py 95 2 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:bie_front -s 3 -c 1 -f -o $P/front python tools/prof_step.py plain_nfs 95 2 > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:att_fold -s 3 -c 1 -f -o $P/fold python tools/prof_step.py plain_nfs 95 2 > /dev/null 2>&1; echo "rc=$?"
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
with G.deterministic():
    for _ in range(3): G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(45, 80))
torch.cuda.synchronize()
PY
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 2 -c 1 -f -o $P/enc python /tmp/enc_prof.py > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 5 -c 1 -f -o $P/vox python /tmp/enc_prof.py > /dev/null 2>&1; echo "rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 8 -c 1 -f -o $P/vox_det python /tmp/enc_prof.py > /dev/null 2>&1; echo "rc=$?"
cat > /tmp/wgrad_prof.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from bmcnet_esr_b200 import kernels as K
from bmcnet_esr_b200 import _lib
b, h, w = 16, 45, 80
rows = b * K.rows_per_image(h, w)
dy = K.pack_nchw(torch.randn(b, 128, h, w, device='cuda'))
x = K.pack_nchw(torch.randn(b, 128, h, w, device='cuda'))
cmap = torch.arange(128, dtype=torch.int32, device='cuda')
gw = torch.zeros(128, 128, 3, 3, device='cuda'); gb = torch.zeros(128, device='cuda')
ws = torch.empty(_lib.lib().bmc_conv_wgrad_workspace_bytes(16, 9, 128), dtype=torch.uint8, device='cuda')
for _ in range(4): K.conv_wgrad(dy, x, 9, b, h, w, cmap, 128, 128, 1.0, gw, gb, ws, 16)
torch.cuda.synchronize()
PY
timeout 600 $NCU --set full --import-source on -k regex:wgrad_tc -s 2 -c 1 -f -o $P/wgrad python /tmp/wgrad_prof.py > /dev/null 2>&1; echo "rc=$?"
for f in slabt_mix slab2_plain3x3 front fold enc vox vox_det wgrad; do python tools/ncu_summary.py $P/$f.ncu-rep > $P/ncu_full_$f.txt 2>&1; done
for f in launches_plain_nfs_B95 launches_bmcnet_nfs_B76 launches_plain_nfs_B1 launches_bmcnet_nfs_B1; do python tools/launch_summary.py $P/$f.csv > $P/$f.txt 2>&1; done
export BMC_B200_LIB=$PWD/bmcnet_esr_b200/libbmc_b200_measure.so
for wl in plain_nfs bmcnet_nfs; do B=95; [ $wl = bmcnet_nfs ] && B=76
BMC_OP_TIMES=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py $wl $B 6 2>&1 | grep -E "optime" > $P/optimes_${wl}_B$B.txt; tail -1 $P/optimes_${wl}_B$B.txt; done
unset BMC_B200_LIB
# sanitizers over the kernels added this round (training kernels, deterministic encoders, the reworked bie_front_tc)
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -x -q -k "wgrad or residual or adam" 2>&1 | tail -6
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoders.py -x -q -k "deterministic_image or goldens" 2>&1 | tail -6
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_model.py -x -q -k "golden or resident or device_resident" 2>&1 | tail -6
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_encoders.py -x -q -k "deterministic_image" 2>&1 | tail -6 ) > $P/sanitizer.txt 2>&1
tail -30 $P/sanitizer.txt
rm -f $P/*.ncu-rep.tmp; ls -la $P/ | head -50
