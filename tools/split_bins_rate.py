"""Rates of the encoders on grids beyond one SM's shared memory: split-bins kernel (role_kernel, default from 2^24 events)
against the global-atomic paths (BMC_ENC_ROLES=0 needs the --measure library).  usage: python tools/split_bins_rate.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bmcnet_esr_b200.dataloader import encodings as G  # noqa: E402


def rate(fn, n, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    return ms, n / ms / 1e6


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(1)
    peak = 6553.0
    ts = torch.sort(torch.rand(n, device=dev, generator=g))[0]
    ps = (torch.randint(0, 2, (n,), device=dev, generator=g) * 2 - 1).float()
    for h, w in ((180, 320), (360, 640)):
        xs = torch.randint(0, w, (n,), device=dev, generator=g).float()
        ys = torch.randint(0, h, (n,), device=dev, generator=g).float()
        for name, fn, bpe in (('channels', lambda: G.events_to_channels(xs, ys, ps, sensor_size=(h, w)), 12),
                              ('voxel B=5', lambda: G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w)), 16),
                              ('voxel_torch B=5', lambda: G.events_to_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(h, w)), 16)):
            ms, gev = rate(fn, n)
            print('%dx%d %-16s %8.3f ms %6.1f Gev/s %5.0f GB/s = %.3f of %d' % (h, w, name, ms, gev, gev * bpe, gev * bpe / peak, peak), flush=True)
        del xs, ys


if __name__ == '__main__':
    main()
