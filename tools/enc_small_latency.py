"""Per-call latency of events_to_channels on one dataloader-sized window (2048 events, 45x80), host-timed over back-to-back
calls (what a per-window caller pays) and device-timed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bmcnet_esr_b200.dataloader import encodings as G
dev = 'cuda'
for n in (2048, 32768, 65536):
    g = torch.Generator(device=dev).manual_seed(1)
    xs = torch.randint(0, 80, (n,), device=dev, generator=g).float(); ys = torch.randint(0, 45, (n,), device=dev, generator=g).float()
    ps = (torch.randint(0, 2, (n,), device=dev, generator=g) * 2 - 1).float()
    for _ in range(20): G.events_to_channels(xs, ys, ps, sensor_size=(45, 80))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(500): out = G.events_to_channels(xs, ys, ps, sensor_size=(45, 80))
    b.record(); torch.cuda.synchronize()
    host = (time.perf_counter() - t0) / 500 * 1e6
    print('n %6d: %.1f us per call (host, back to back), %.1f us (device timeline)' % (n, host, a.elapsed_time(b) * 1e3 / 500))
