"""Where a batch-1 frame's time goes: device-resident step() back to back (GPU-bound) vs forward() with a per-frame
synchronisation (the infer_BMCNet.py:54-68 pattern).  python tools/lat_parts.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle.make_golden import synth_counts
from bmcnet_esr_b200.models.BMCNet import BMCNet
from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
dev = torch.device('cuda', 0)
for kind in ('plain', 'full'):
    sd, _ = bench.load_state(kind)
    m = (BMCNet_plain if kind == 'plain' else BMCNet)(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    h, w = 45, 80
    xs = [synth_counts(1, h, w, 50 + i).to(dev) for i in range(8)]
    with torch.no_grad():
        for i in range(10):
            m.step(xs[i % 8], reset=(i == 0))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 200
        for i in range(n):
            m.step(xs[i % 8], reset=False)
        b.record()
        torch.cuda.synchronize()
        print(kind, 'step() back to back: %.4f ms/frame (%d graph kernels)' % (a.elapsed_time(b) / n, m._engine.launches_per_step))
        n_state = 2 if kind == 'plain' else 4
        st = [torch.zeros(1, 128, h, w, device=dev) for _ in range(n_state - 1)] + [torch.zeros(1, 32, h, w, device=dev)]
        st = list(m(xs[0], *st, True))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            st = list(m(xs[i % 8], *st, False))
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(kind, 'forward() host enqueue: %.4f ms/frame; until drained %.4f ms/frame' % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
