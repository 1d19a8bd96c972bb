"""Find the first call that invalidates a CUDA-graph capture of the training iteration."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench
from bmcnet_esr_b200 import _lib, kernels as K
from bmcnet_esr_b200.models import _train as T
from bmcnet_esr_b200.models.BMCNet import BMCNet
from oracle.make_golden import synth_counts

b, h, w, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device('cuda', 0)
sd, _ = bench.load_state('full')
m = BMCNet(4, 128, 5); m.load_state_dict(sd, strict=True); m = m.to(dev).train()
opt = T.FusedAdamAMSGrad(m.parameters())
xs = [synth_counts(b, h, w, 3000 + s).to(dev) for s in range(steps)]
gts = [torch.rand(b, 2, 4 * h, 4 * w, device=dev) for _ in range(steps)]

def probe(tag):
    try:
        torch.cuda.is_current_stream_capturing()
    except Exception as e:
        print('CAPTURE INVALID after', tag, '::', str(e).splitlines()[0]); traceback.print_stack(limit=8); os._exit(3)

for name in ('conv_gemm', 'relu_backward', 'conv_wgrad', 'layernorm_rows', 'layernorm_rows_backward'):
    orig = getattr(K, name)
    def wrap(*a, _o=orig, _n=name, **k):
        probe('before ' + _n)
        r = _o(*a, **k)
        probe(_n)
        return r
    setattr(K, name, wrap)

def body():
    opt.zero_grad(); probe('zero_grad')
    st = [torch.zeros(b, 128, h, w, device=dev) for _ in range(3)] + [torch.zeros(b, 32, h, w, device=dev)]
    loss, init = 0, True
    for x, gt in zip(xs, gts):
        st = list(m(x, *st, init)); init = False; probe('forward')
        loss = loss + F.mse_loss(st[-1], gt); probe('mse')
    loss.backward(); probe('backward')
    return loss.detach()

for _ in range(2):
    body(); opt.step()
torch.cuda.synchronize()
orig_body = T.GraphedIteration._body
def pb(self):
    probe('enter body')
    r = orig_body(self)
    probe('exit body')
    return r
T.GraphedIteration._body = pb
orig_bie = T.bie
def bie_p(*a, **k):
    r = orig_bie(*a, **k); probe('bie'); return r
T.bie = bie_p
it = T.GraphedIteration(m, opt, xs, gts, warmup=1)
print('captured OK', it().item())
