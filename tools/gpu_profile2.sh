#!/bin/bash
# late-round additions to the ncu evidence: conv_slab2 at the BMCNet bench shape (4 jobs, B=76), the stack encoder, the sweep
mkdir -p gpurun_out/prof
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:conv_slab2 -s 1 -c 1 -f -o gpurun_out/prof/slab2_bmcnet3x3 python tools/time_conv.py 76 4 2 > /dev/null 2>&1; echo "rc=$?"
cat > /tmp/stack_prof.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from bmcnet_esr_b200.dataloader import encodings as G
n = 100_000_000
xs = torch.rand(n, device='cuda') * 80; ys = torch.rand(n, device='cuda') * 45
ps = (torch.rand(n, device='cuda') < 0.5).float() * 2 - 1
ts = torch.sort(torch.rand(n, device='cuda'))[0]
for _ in range(3): G.events_to_stack_polarity(xs, ys, ts, ps, 5, sensor_size=(45, 80))
torch.cuda.synchronize()
PY
timeout 600 $NCU --set full --import-source on -k regex:scatter_kernel -s 2 -c 1 -f -o gpurun_out/prof/stack python /tmp/stack_prof.py > /dev/null 2>&1; echo "rc=$?"
for f in slab2_bmcnet3x3 stack; do python tools/ncu_summary.py gpurun_out/prof/$f.ncu-rep > gpurun_out/prof/ncu_full_$f.txt 2>&1; done
timeout 600 python tools/enc_sweep.py > gpurun_out/enc_sweep.md 2> gpurun_out/enc_sweep.err; echo "sweep rc=$?"
grep -i -E "dram__bytes|duration|tensor" gpurun_out/prof/ncu_full_slab2_bmcnet3x3.txt | head -12
grep -i -E "dram__bytes|duration|lsu|inst_exec|issue" gpurun_out/prof/ncu_full_stack.txt | head -12
rm -f gpurun_out/prof/*.ncu-rep.tmp
