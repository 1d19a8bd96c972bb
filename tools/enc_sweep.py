"""BASELINE configs[2]: the event-encoding sweep -- 1e6 .. 1e9 synthetic events into per-polarity count grids,
voxel grids and polarity stacks at the three grids SURVEY.md 8(d) names, one GPU, device-timed (CUDA events, warm,
median of 5).  Every count grid is also checked bit-exactly against torch.bincount on the same events (an
independent device-side histogram; the CPU oracle covers <= 1e8 in tests/) and every stack against its bin
structure.  Writes a markdown table to stdout.      python tools/enc_sweep.py [max_events=1e9]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bmcnet_esr_b200.dataloader import encodings as G     # noqa: E402

dev = 'cuda'
n_max = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
HBM = 6553.3


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2] * 1e-3, out


def bincount_channels(xs, ys, ps, h, w):
    """per-polarity counts the library way: flat pixel index of in-range events, y flipped (encodings.py:256-258)"""
    out = torch.zeros(2, h * w, device=dev)
    step = 250_000_000
    for lo in range(0, len(xs), step):
        x, y, p = xs[lo:lo + step].long(), ys[lo:lo + step].long(), ps[lo:lo + step]
        pix = (h - 1 - y) * w + x
        for c, m in enumerate((p > 0, p < 0)):
            out[c] += torch.bincount(pix[m], minlength=h * w).float()
    return out.view(2, h, w)


print('| events | grid | encoder | ms | Gevents/s | algorithmic GB/s | of %.0f GB/s | check |' % HBM)
print('|---|---|---|---|---|---|---|---|')
for n in (1_000_000, 10_000_000, 100_000_000, 1_000_000_000):
    if n > n_max:
        break
    for (h, w) in ((45, 80), (180, 320), (360, 640)):
        g = torch.Generator(device=dev).manual_seed(n % 1000 + h)
        xs = torch.floor(torch.rand(n, device=dev, generator=g) * w)
        ys = torch.floor(torch.rand(n, device=dev, generator=g) * h)
        ps = (torch.rand(n, device=dev, generator=g) < 0.5).float() * 2 - 1
        if n <= 100_000_000:
            ts = torch.sort(torch.rand(n, device=dev, generator=g))[0]
            ts = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)
        else:       # 1e9: evenly spaced stamps (float32 leaves ~60 equal neighbours each, as a real recording this long would)
            ts = torch.linspace(0, 1, n, device=dev)
            ts = ts / (ts[-1] + 1e-6)
        rows = []
        s, cnt = timed(lambda: G.events_to_channels(xs, ys, ps, sensor_size=(h, w)))
        ok = torch.equal(cnt, bincount_channels(xs, ys, ps, h, w))
        rows.append(('events_to_channels', 12, s, 'bit-exact vs bincount' if ok else 'MISMATCH'))
        s, vox = timed(lambda: G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w)))
        mass = abs(float(vox.sum(dtype=torch.float64)) - float(ps.sum(dtype=torch.float64)))
        rows.append(('events_to_voxel B=5', 16, s, 'mass error %.1e of n' % (mass / n)))
        s, st = timed(lambda: G.events_to_stack_polarity(xs, ys, ts, ps, 5, sensor_size=(h, w)))
        tot = float(st.sum(dtype=torch.float64))
        # every event in >= 1 bin, boundary events (and their equal-stamp neighbours) in two; planes are counts
        okst = n <= tot and float(st.min()) >= 0 and torch.equal(st.sum(1).clamp(max=0), torch.zeros_like(st[:, 0]))
        dup = tot - n
        rows.append(('events_to_stack_polarity B=5', 12, s, ('counts >= 0, %d boundary double counts' % dup) if okst else 'MISMATCH'))
        if n <= 100_000_000:
            nw = n // 2048
            offs = torch.arange(nw + 1, device=dev, dtype=torch.int64) * 2048
            if 2 * h * w * 4 <= 200_000:      # the one-CTA-per-window kernel keeps a window grid in shared memory
                s, wins = timed(lambda: G.events_to_channels_windows(xs, ys, ps, offs, sensor_size=(h, w)))
                k = min(nw, 3)
                okw = all(torch.equal(wins[i], bincount_channels(xs[i * 2048:(i + 1) * 2048], ys[i * 2048:(i + 1) * 2048],
                                                                 ps[i * 2048:(i + 1) * 2048], h, w)) for i in range(k))
                okw = okw and float(wins.sum(dtype=torch.float64)) == nw * 2048
                rows.append(('channels, %d windows of 2048' % nw, 12 + 2 * h * w * 4 / 2048.0, s,
                             'bit-exact vs bincount (first windows), total exact' if okw else 'MISMATCH'))
                del wins
        for name, bpe, s, chk in rows:
            print('| %.0e | %dx%d | %s | %.3f | %.1f | %.0f | %.2f | %s |' % (n, h, w, name, s * 1e3, n / s / 1e9, n * bpe / s / 1e9,
                                                                       n * bpe / s / 1e9 / HBM, chk), flush=True)
        del xs, ys, ps, ts, cnt, vox, st
        torch.cuda.empty_cache()
