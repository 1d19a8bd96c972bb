"""Graphed training iteration: run-to-run reproducibility, and wgrad side stream on / off agreement (losses, gradients)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bmcnet_esr_b200.models.BMCNet import BMCNet
from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration
from oracle.make_golden import synth_counts

b, h, w, steps = 2, 22, 40, 3
dev = torch.device('cuda', 0)
sd, _ = bench.load_state('full')
xs = [synth_counts(b, h, w, 900 + s).to(dev) for s in range(steps)]
g = torch.Generator().manual_seed(3)
gts = [torch.poisson(torch.full((b, 2, 4 * h, 4 * w), 0.3), generator=g).to(dev) for _ in range(steps)]

def run(branch):
    m = BMCNet(4, 128, 5); m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True); m = m.to(dev).train()
    opt = FusedAdamAMSGrad(m.parameters(), lr=1e-4)
    it = GraphedIteration(m, opt, xs, gts, wgrad_stream=branch)
    out = []
    for _ in range(3):
        loss = it().item()
        out.append((loss, opt.grad.clone()))
    sizes = [(n, p.numel()) for n, p in m.named_parameters()]
    return out, sizes

(a, sizes), (a2, _), (c, _) = run(True), run(True), run(False)
for i in range(3):
    ga, ga2, gc = a[i][1], a2[i][1], c[i][1]
    print('iter', i, 'loss multi %.9f multi2 %.9f single %.9f' % (a[i][0], a2[i][0], c[i][0]),
          '| grad multi vs multi2: max %.3e equal %s | multi vs single: max %.3e rel-to-max %.3e' % (
              (ga - ga2).abs().max().item(), torch.equal(ga, ga2), (ga - gc).abs().max().item(), ((ga - gc).abs().max() / gc.abs().max()).item()))

off = 0
ga, ga2, gc = a[0][1], a2[0][1], c[0][1]
for n, k in sizes:
    e1 = (ga[off:off + k] - gc[off:off + k]).abs().max().item(); e2 = (ga[off:off + k] - ga2[off:off + k]).abs().max().item()
    mx = gc[off:off + k].abs().max().item()
    if e1 > 1e-3 * mx or e2 > 0:
        print('%-50s max %.3e  multi-single %.3e (%.1f%%)  multi-multi2 %.3e' % (n, mx, e1, 100 * e1 / max(mx, 1e-30), e2))
    off += k
