#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_encoders.py -x -q 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from bmcnet_esr_b200.dataloader import encodings as G
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for n in (10**7, 10**8, 4 * 10**8):
    xs = torch.rand(n, device='cuda') * 80; ys = torch.rand(n, device='cuda') * 45
    ps = (torch.rand(n, device='cuda') < 0.5).float() * 2 - 1
    ts = torch.sort(torch.rand(n, device='cuda'))[0]
    for B, (h, w) in ((5, (45, 80)), (5, (180, 320))):
        sx = 1.0 if h == 45 else 4.0
        ms = t(lambda: G.events_to_voxel(xs * sx, ys * sx, ts, ps, B, sensor_size=(h, w)))
        base = t(lambda: (xs * sx, ys * sx))
        print('voxel n=%.0e B=%d %dx%d: %.3f ms (minus scaling %.3f) -> %.0f GB/s' % (n, B, h, w, ms, base, 16 * n / (ms - base) / 1e6))
    ms = t(lambda: G.events_to_channels(xs, ys, ps, sensor_size=(45, 80)))
    print('channels n=%.0e: %.3f ms -> %.0f GB/s' % (n, ms, 12 * n / ms / 1e6))
    ms = t(lambda: G.events_to_stack_polarity(xs, ys, ts, ps, 5, sensor_size=(45, 80)))
    print('stack_polarity n=%.0e: %.3f ms -> %.0f GB/s' % (n, ms, 12 * n / ms / 1e6))
    del xs, ys, ps, ts
PY
