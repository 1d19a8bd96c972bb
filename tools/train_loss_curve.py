"""Loss per iteration of the bench's training workload (surrogate weights, synthetic targets), for the data of every rank of an
8-GPU run, one GPU, no averaging: is any rank's sequence the source of a non-finite gradient?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bmcnet_esr_b200.models.BMCNet import BMCNet
from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration
from oracle.make_golden import synth_counts
b, h, w, steps = 2, 45, 80, 8
dev = torch.device('cuda', 0)
sd, _ = bench.load_state('full')
for rank in ([int(a) for a in sys.argv[1:]] or range(8)):
    m = BMCNet(4, 128, 5); m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True); m = m.to(dev).train()
    opt = FusedAdamAMSGrad(m.parameters(), lr=float(os.environ.get('LR', '1e-4')))
    xs = [synth_counts(b, h, w, 3000 + 17 * rank + s).to(dev) for s in range(steps)]
    gts = bench.train_targets(xs, rank)
    it = GraphedIteration(m, opt, xs, gts, warmup=1)
    out = []
    for _ in range(14):
        l = it().item()
        out.append('%.4g/%s' % (l, 'ok' if bool(torch.isfinite(opt.grad).all()) else 'NONFINITE-GRAD'))
    print('rank', rank, ' '.join(out), '| overflows', it.overflows, 'loss_scale', m.loss_scale, flush=True)
    del it, m, opt
    torch.cuda.empty_cache()
