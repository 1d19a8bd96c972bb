"""Kernel histogram of one replay of the graphed BMCNet training iteration (torch.profiler / CUPTI).
usage: python tools/train_profile.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402
from bmcnet_esr_b200.models.BMCNet import BMCNet  # noqa: E402
from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration  # noqa: E402
from oracle.make_golden import synth_counts  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    h, w, steps = 45, 80, 8
    dev = torch.device('cuda', 0)
    sd, _ = bench.load_state('full')
    m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    opt = FusedAdamAMSGrad(m.parameters(), lr=bench.TRAIN_LR)
    xs = [synth_counts(b, h, w, 3000 + s).to(dev) for s in range(steps)]
    gts = bench.train_targets(xs, 0)
    it = GraphedIteration(m, opt, xs, gts, warmup=1)
    it(); it()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        it()
        torch.cuda.synchronize()
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            n = e.name[:90]
            c = rows.setdefault(n, [0, 0.0])
            c[0] += 1
            c[1] += e.device_time
    # (CUPTI may hand back the records of more than one replay: normalise by the Adam launches, one per iteration)
    reps = max(1, sum(v[0] for k, v in rows.items() if 'adam_amsgrad' in k))
    tot = sum(v[1] for v in rows.values()) / reps
    cnt = sum(v[0] for v in rows.values()) // reps
    print('overflows', it.overflows)
    print('batch %d: %d kernels, %.1f ms of kernel time per iteration (%d iteration(s) in the trace)' % (b, cnt, tot / 1e3, reps))
    for n, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:45]:
        print('%7d %9.1f us %5.1f%% %6.2f us/launch  %s' % (c // reps, t / reps, 100 * t / reps / tot, t / c, n))


if __name__ == '__main__':
    main()
