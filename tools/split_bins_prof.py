import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bmcnet_esr_b200.dataloader import encodings as G
n = 100_000_000; dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
ts = torch.sort(torch.rand(n, device=dev, generator=g))[0]
ps = (torch.randint(0, 2, (n,), device=dev, generator=g) * 2 - 1).float()
which = sys.argv[1] if len(sys.argv) > 1 else 'voxel'
if which == 'voxel':
    h, w = 180, 320
    xs = torch.randint(0, w, (n,), device=dev, generator=g).float(); ys = torch.randint(0, h, (n,), device=dev, generator=g).float()
    for _ in range(3): G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w))
else:
    h, w = 360, 640
    xs = torch.randint(0, w, (n,), device=dev, generator=g).float(); ys = torch.randint(0, h, (n,), device=dev, generator=g).float()
    for _ in range(3): G.events_to_channels(xs, ys, ps, sensor_size=(h, w))
torch.cuda.synchronize()
