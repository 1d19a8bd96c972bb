import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bmcnet_esr_b200.dataloader import encodings as G
dev = 'cuda'
h, w, B = 180, 320, 5
n = int(sys.argv[1]) if len(sys.argv) > 1 else 301_056
mode = sys.argv[2] if len(sys.argv) > 2 else 'struct'
g = torch.Generator(device=dev).manual_seed(1)
ts = torch.sort(torch.rand(n, device=dev, generator=g))[0]
i = torch.arange(n, device=dev)
xs = (i % w).float(); ys = ((i // w) % h).float(); ps = torch.ones(n, device=dev)
if 'randxy' in mode:
    xs = torch.randint(0, w, (n,), device=dev, generator=g).float(); ys = torch.randint(0, h, (n,), device=dev, generator=g).float()
if 'randp' in mode:
    ps = (torch.randint(0, 2, (n,), device=dev, generator=g) * 2 - 1).float()
G.SPLIT_BINS = False
ref = G.events_to_voxel(xs, ys, ts, ps, B, sensor_size=(h, w))
G.SPLIT_BINS = True
got = G.events_to_voxel(xs, ys, ts, ps, B, sensor_size=(h, w))
d = (got - ref)
# per event expected contributions
tn = ts * (B - 1); fl = tn.floor(); fr = tn - fl
pix = (h - 1 - ys.long()) * w + xs.long()
dl = d.view(B, -1)
lost = []
for s in range(B - 1):
    sel = (fl == s).nonzero().flatten()
    e_lo = dl[s, pix[sel]]; e_hi = dl[s + 1, pix[sel]]
    # event lost if error in bin s ~ -(1-fr) and bin s+1 ~ -fr ; duplicated if +
    for sign, name in ((-1, 'lost'), (1, 'dup')):
        m = ((e_lo - sign * ps[sel] * (1 - fr[sel])).abs() < 1e-3) & ((e_hi - sign * ps[sel] * fr[sel]).abs() < 1e-3) & ((1 - fr[sel]).abs() + fr[sel].abs() > 0)
        m &= (e_lo.abs() + e_hi.abs()) > 1e-3
        idx = sel[m]
        lost.append((s, name, idx))
tot_bad = int((d.abs() > 1e-4).sum())
print('n', n, mode, 'bad bins', tot_bad)
for s, name, idx in lost:
    if len(idx):
        print('slot', s, name, len(idx), 'first', idx[:12].tolist(), 'last', idx[-5:].tolist())
        per3, per4 = ((n + 48) // 49 + 3) // 4 * 4, ((n + 36) // 37 + 3) // 4 * 4
        for T, K, per in ((3, 2, per3), (4, 3, per4)):
            E = 1024 * K
            ii = idx[:8]
            print('   T=%d: cluster %s off %s rank %s j %s tid %s' % (T, (ii // per).tolist(), (ii % per).tolist(), ((ii % per) // E).tolist(), (((ii % per) % E) // 1024).tolist(), ((ii % per) % 1024).tolist()))
b = (fl[1:] != fl[:-1]).nonzero().flatten() + 1
print('slot boundaries at event', b.tolist())
