"""Voxel / count encoder rates at 45x80 (and 180x320), 4e8 events: python tools/vox_rate.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bmcnet_esr_b200.dataloader import encodings as G
n = 400_000_000
dev = 'cuda'
for (h, w) in ((45, 80), (180, 320)):
    xs = torch.rand(n, device=dev) * w; ys = torch.rand(n, device=dev) * h
    ps = (torch.rand(n, device=dev) < 0.5).float() * 2 - 1
    ts = torch.sort(torch.rand(n, device=dev))[0]
    def timed(fn, reps=3):
        fn(); fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    ms = timed(lambda: G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w)))
    print('%dx%d voxel B=5      %.3f ms %.1f Gev/s %.0f GB/s = %.3f of 6553' % (h, w, ms, n / ms / 1e6, 16 * n / ms / 1e6, 16 * n / ms / 1e6 / 6553.3), flush=True)
    with G.deterministic():
        ms = timed(lambda: G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w)))
    print('%dx%d voxel B=5 det  %.3f ms %.1f Gev/s' % (h, w, ms, n / ms / 1e6), flush=True)
    ms = timed(lambda: G.events_to_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(h, w)))
    print('%dx%d voxel_torch    %.3f ms %.1f Gev/s' % (h, w, ms, n / ms / 1e6), flush=True)
    ms = timed(lambda: G.events_to_channels(xs, ys, ps, sensor_size=(h, w)))
    print('%dx%d channels       %.3f ms %.1f Gev/s' % (h, w, ms, n / ms / 1e6), flush=True)
    del xs, ys, ps, ts
