"""Where a stack-encoder call spends its time: host early-out check, the C entry alone, the whole Python call."""
import os, sys, time, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bmcnet_esr_b200 import _lib
from bmcnet_esr_b200.dataloader import encodings as G
dev = 'cuda'; h, w, B = 45, 80, 5
def timed(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6
for n in (1_000_000, 100_000_000, 400_000_000):
    xs = torch.rand(n, device=dev) * w; ys = torch.rand(n, device=dev) * h
    ps = (torch.rand(n, device=dev) < 0.5).float() * 2 - 1
    ts = torch.sort(torch.rand(n, device=dev))[0]
    out = torch.empty(2, B, h, w, device=dev)
    nbytes = _lib.lib().bmc_encode_workspace_bytes(out.numel())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    centry = lambda: _lib.check(_lib.lib().bmc_encode_stack(p(xs), p(ys), p(ts), p(ps), n, B, h, w, 1, p(out), p(ws), nbytes, _lib.ENC_MUTATE, _lib.stream_ptr()))
    print('n=%.0e  early_out check: %.0f us (wall %.0f)   C entry: %.0f us (wall %.0f)   python call: %.0f us (wall %.0f)   channels: %.0f us' % (
        (n,) + timed(lambda: G._early_out(ts, B, (h, w), dev)) + timed(centry) + timed(lambda: G.events_to_stack_polarity(xs, ys, ts, ps, B, sensor_size=(h, w)))
        + (timed(lambda: G.events_to_channels(xs, ys, ps, sensor_size=(h, w)))[0],)), flush=True)
    del xs, ys, ps, ts
