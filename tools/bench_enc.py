"""Encoder kernel timings on one GPU (CUDA events, warm): Gevents/s and GB/s of algorithmic traffic.
    python tools/bench_enc.py [n_events]        (BMC_VOXEL_MODE=0|1|2 selects the voxel kernel form)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bmcnet_esr_b200.dataloader import encodings as G

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
only = sys.argv[2] if len(sys.argv) > 2 else ''
dev = 'cuda'
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for (h, w) in ((45, 80), (180, 320)):
    xs = torch.rand(n, device=dev) * w; ys = torch.rand(n, device=dev) * h
    ps = (torch.rand(n, device=dev) < 0.5).float() * 2 - 1
    ts = torch.sort(torch.rand(n, device=dev))[0]
    tu = torch.rand(n, device=dev)
    rows = [('channels', 12, lambda: G.events_to_channels(xs, ys, ps, sensor_size=(h, w))),
            ('voxel B=5 sorted', 16, lambda: G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w))),
            ('voxel B=5 unsorted', 16, lambda: G.events_to_voxel(xs, ys, tu, ps, 5, sensor_size=(h, w))),
            ('voxel_torch B=5', 16, lambda: G.events_to_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(h, w))),
            ('stack_polarity B=5', 12, lambda: G.events_to_stack_polarity(xs, ys, ts, ps, 5, sensor_size=(h, w)))]
    for name, bpe, fn in rows:
        if only not in name:
            continue
        s = timed(fn)
        print('enc %-20s %3dx%-3d n=%.0e mode=%s : %7.1f Gev/s %7.0f GB/s' % (name, h, w, n, os.environ.get('BMC_VOXEL_MODE', '-'), n / s / 1e9, n * bpe / s / 1e9), flush=True)
    del xs, ys, ps, ts, tu
