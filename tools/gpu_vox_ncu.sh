#!/bin/bash
# re-capture of the voxel encoder kernels after the fixed-point shared-memory path (kSmemFix)
mkdir -p gpurun_out/prof
P=gpurun_out/prof
cat > /tmp/enc_prof.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from bmcnet_esr_b200.dataloader import encodings as G
n = 100_000_000
xs = torch.rand(n, device='cuda') * 80; ys = torch.rand(n, device='cuda') * 45
ps = (torch.rand(n, device='cuda') < 0.5).float() * 2 - 1
ts = torch.sort(torch.rand(n, device='cuda'))[0]
for _ in range(3): G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(45, 80))
torch.cuda.synchronize()
PY
timeout 600 ncu --clock-control none --set full --import-source on -k regex:scatter_kernel -s 2 -c 1 -f -o $P/vox python /tmp/enc_prof.py > /dev/null 2>&1; echo "rc=$?"
python tools/ncu_summary.py $P/vox.ncu-rep > $P/ncu_full_vox.txt 2>&1
cat $P/ncu_full_vox.txt
python tools/vox_rate.py
