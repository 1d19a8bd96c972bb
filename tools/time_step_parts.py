"""Where a device-resident step spends its time: encode / pack+graph+emit, measured with CUDA events.
    python tools/time_step_parts.py workload [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bmcnet_esr_b200.dataloader import encodings as G
from bmcnet_esr_b200.models.BMCNet import BMCNet
from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain

wl = sys.argv[1]
kind, h, w, n_win, _ = bench.WORKLOADS[wl]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 95
dev = torch.device('cuda', 0)
sd, _ = bench.load_state(kind)
model = (BMCNet_plain if kind == 'plain' else BMCNet)(4, 128, 5)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
steps = 60
ev = bench.synth_stream(steps, B, n_win, h, w, 1, dev)
offsets = torch.arange(0, B * 2 * n_win + 1, n_win, dtype=torch.int64, device=dev)
def timed(fn, n):
    for k in range(5): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(n): fn(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
def enc(k):
    return G.events_to_channels_windows(ev[k][0], ev[k][1], ev[k][2], offsets, sensor_size=(h, w))
cnt = enc(0)
x = cnt.view(B, 2, 2, h, w).transpose(1, 2)
model.step(x, reset=True)
t_enc = timed(lambda k: enc(k % steps), steps)
t_step = timed(lambda k: model.step(x, reset=False), steps)
t_both = timed(lambda k: model.step(enc(k % steps).view(B, 2, 2, h, w).transpose(1, 2), reset=False), steps)
print('parts %s B=%d: encode %.1f us, step (pack + graph + emit) %.1f us, both %.1f us' % (wl, B, t_enc, t_step, t_both))
