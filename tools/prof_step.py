"""Run a few device-resident recurrent steps of one bench workload (for ncu launch lists).
python tools/prof_step.py workload [batch] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bmcnet_esr_b200.dataloader import encodings as G
from bmcnet_esr_b200.models.BMCNet import BMCNet
from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain

wl = sys.argv[1]
kind, h, w, n_win, _ = bench.WORKLOADS[wl]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 19
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device('cuda', 0)
sd, _ = bench.load_state(kind)
model = (BMCNet_plain if kind == 'plain' else BMCNet)(4, 128, 5)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
ev = bench.synth_stream(steps, B, n_win, h, w, 1, dev)
offsets = torch.arange(0, B * 2 * n_win + 1, n_win, dtype=torch.int64, device=dev)
for k in range(steps):
    cnt = G.events_to_channels_windows(ev[k][0], ev[k][1], ev[k][2], offsets, sensor_size=(h, w))
    torch.cuda.nvtx.range_push('step%d' % k)
    model.step(cnt.view(B, 2, 2, h, w).transpose(1, 2), reset=(k == 0))
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print('done', model._engine.launches_per_step)
