"""GPU parity of the training step (reference train.py:202-237, config/train_nfs.yml:28-34):
the backward kernels (dgrad = bmc_conv_gemm on mirrored / transposed weights, wgrad = bmc_conv_wgrad, ReLU'),
the autograd path of the drop-in modules in train() mode, and the fused Adam(amsgrad) step -- against
torch.autograd on the fp32 oracle (oracle/train_step.py, pinned to the reference's modules + MSELoss + Adam).

Bars (fp16 operands / activations / activation gradients under a static loss scale, fp32 accumulation and fp32
weight gradients): per-kernel gradients within 2e-3 of the tensor's max-abs; whole-sequence gradients of every
unique parameter within 1.5 % of that parameter's max-abs gradient and 1 % in relative L2 (measured: 0.8 % / 0.6 %); sequence loss within 1e-3
relative; Adam update bit-close (1e-6) to the restated update rule."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bmcnet_fp32 as O
from oracle import train_step as T
from oracle.make_golden import synth_counts

pytestmark = pytest.mark.gpu

# the fp32 references below must be fp32: cuDNN convolutions default to TF32 (10-bit mantissa) on this GPU
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _lib_act():
    from bmcnet_esr_b200 import _lib
    return str(_lib.act_dtype()).replace('torch.', '')


def _rel_max(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _rel_l2(a, b):
    return (a - b).norm().item() / max(b.norm().item(), 1e-30)


@pytest.mark.parametrize('taps,x_ch,cin,n_out', [(9, 128, 128, 128), (1, 128, 128, 128), (9, 64, 22, 128), (9, 128, 128, 32)])
def test_wgrad_and_dgrad_kernels_vs_autograd(taps, x_ch, cin, n_out):
    """One convolution: dX (dgrad) and dW, db (wgrad) against torch.autograd of F.conv2d in fp32 on the same
    fp16-rounded operands."""
    from bmcnet_esr_b200 import kernels as K
    from bmcnet_esr_b200.models import _train as TR
    torch.manual_seed(taps * 1000 + x_ch + n_out)
    b, h, w = 3, 13, 21
    k = 3 if taps == 9 else 1
    conv = torch.nn.Conv2d(cin, n_out, k, 1, k // 2).cuda()
    with torch.no_grad():
        conv.weight.copy_(conv.weight.half().float())
    x = (torch.randn(b, cin, h, w, device='cuda') * 0.7).half().float().requires_grad_(True)
    tc = TR._Ctx(1.0)
    seg = list(range(cin)) + [-1] * (x_ch - cin)
    xp = TR.to_packed(x, x_ch)
    out = TR.conv(tc, conv, [xp], [seg], (b, h, w), relu=False)     # ReLU' is covered by the ResidualBlock test below
    y = TR.from_packed(out, b, n_out, h, w)
    gy = (torch.randn_like(y) * 0.5).half().float()
    y.backward(gy)
    # reference
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=k // 2)
    yr.backward(gy)
    assert _rel_max(y.detach(), yr.detach()) <= 2e-3
    assert _rel_max(x.grad, xr.grad) <= 2e-3, ('dgrad', _rel_max(x.grad, xr.grad))
    assert _rel_max(conv.weight.grad, wr.grad) <= 2e-3, ('wgrad', _rel_max(conv.weight.grad, wr.grad))
    assert _rel_max(conv.bias.grad, br.grad) <= 2e-3, ('bgrad', _rel_max(conv.bias.grad, br.grad))


def test_residual_block_backward_end_to_end():
    """ResidualBlock_noBN (submodules.py:31-35): x + conv2(relu(conv1(x))), gradients of both convs and of x."""
    from bmcnet_esr_b200.models import _train as TR
    from bmcnet_esr_b200.models.submodules import ResidualBlock_noBN
    torch.manual_seed(3)
    b, h, w = 2, 31, 56
    blk = ResidualBlock_noBN(128).cuda()
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_((p * 4 + 0.01).half().float())
    x = torch.randn(b, 128, h, w, device='cuda').half().float().requires_grad_(True)
    tc = TR._Ctx(64.0)
    xs = TR._ScaleGrad.apply(x, 1.0 / 64.0)
    y = TR._ScaleGrad.apply(TR.from_packed(TR.resblock(tc, blk, TR.to_packed(xs, 128), (b, h, w)), b, 128, h, w), 64.0)
    gy = torch.randn_like(y).half().float()
    y.backward(gy)
    xr = x.detach().clone().requires_grad_(True)
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    yr = xr + F.conv2d(F.relu(F.conv2d(xr, ps['conv1.weight'], ps['conv1.bias'], padding=1)), ps['conv2.weight'], ps['conv2.bias'], padding=1)
    yr.backward(gy)
    assert _rel_max(y.detach(), yr.detach()) <= 2e-3
    # a pre-activation within fp16 rounding of zero flips its ReLU mask and moves single gradient entries by a whole
    # term, so the max-norm bar is looser than the L2 one here
    assert _rel_l2(x.grad, xr.grad) <= 3e-3 and _rel_max(x.grad, xr.grad) <= 2e-2, (_rel_l2(x.grad, xr.grad), _rel_max(x.grad, xr.grad))
    for n, p in blk.named_parameters():
        assert _rel_l2(p.grad, ps[n].grad) <= 3e-3, (n, _rel_l2(p.grad, ps[n].grad))
        assert _rel_max(p.grad, ps[n].grad) <= 1e-2, (n, _rel_max(p.grad, ps[n].grad))


def test_layernorm_backward_kernel_vs_autograd():
    """LayerNormFunction.backward (submodules.py:142-154): dx, dgamma, dbeta of the channel LayerNorm kernel pair against
    torch.autograd on the fp32 formula, on fp16-rounded inputs (14,000 rows incl. rows of all zeros like the halo)."""
    from bmcnet_esr_b200.models import _train as TR
    from bmcnet_esr_b200.models.submodules import LayerNorm2d
    torch.manual_seed(5)
    rows = 14000
    norm = LayerNorm2d(128).cuda()
    with torch.no_grad():
        norm.weight.copy_(0.87 + 0.05 * torch.randn(128))
        norm.bias.copy_(0.03 * torch.randn(128))
    x = (torch.randn(rows, 128, device='cuda') * 1.5 + 0.3).half()
    x[::97] = 0
    xg = x.clone().requires_grad_(True)
    tc = TR._Ctx(8.0)
    y = TR.layernorm_rows(tc, norm, xg)
    gy = (torch.randn(rows, 128, device='cuda') * 0.5).half()
    y.backward(gy)
    xr = x.float().requires_grad_(True)
    g, b = norm.weight.detach().clone().requires_grad_(True), norm.bias.detach().clone().requires_grad_(True)
    mu = xr.mean(1, keepdim=True)
    var = ((xr - mu) ** 2).mean(1, keepdim=True)
    yr = g * ((xr - mu) / (var + norm.eps).sqrt()) + b
    yr.backward(gy.float())
    assert _rel_max(y.detach().float(), yr.detach()) <= 2e-3
    assert _rel_l2(xg.grad.float(), xr.grad) <= 2e-3, _rel_l2(xg.grad.float(), xr.grad)
    # (the loss scale of the context divides the parameter gradients: 8.0)
    assert _rel_max(norm.weight.grad * 8.0, g.grad) <= 2e-3 and _rel_max(norm.bias.grad * 8.0, b.grad) <= 2e-3


def test_adam_amsgrad_kernel_matches_update_rule():
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad
    torch.manual_seed(0)
    p0 = {'a': torch.randn(7, 5), 'b': torch.randn(11), 'c': torch.randn(128, 128, 3, 3) * 0.05}
    mine = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in p0.items()}
    opt = FusedAdamAMSGrad(list(mine.values()), lr=1e-3, weight_decay=1e-2)
    ref, state = {k: v.clone() for k, v in p0.items()}, {}
    for step in range(5):
        grads = {k: torch.randn_like(v) * (0.1 if step == 3 else 1.0) for k, v in p0.items()}      # a small step exercises vmax
        opt.zero_grad()
        for k in mine:
            mine[k].grad.copy_(grads[k])
        v0 = mine['a']._version
        opt.step()
        assert mine['a']._version > v0                      # the raw-pointer update is visible to version checks
        T.adam_amsgrad_step(ref, grads, state, lr=1e-3, weight_decay=1e-2)
    for k in ref:
        assert torch.allclose(mine[k].detach().cpu(), ref[k], rtol=2e-6, atol=1e-7), k


def _sequence(m, xs, gts, n_state):
    b, _, _, h, w = xs[0].shape
    st = [torch.zeros(b, 128, h, w, device='cuda') for _ in range(n_state)] + [torch.zeros(b, 32, h, w, device='cuda')]
    loss, init = 0, True
    for x, gt in zip(xs, gts):                                   # train.py:206-234
        st = list(m(x.cuda().transpose(1, 2).contiguous().transpose(1, 2), *st, init))
        init = False
        pred = st[-1]
        if pred.shape[-2:] != gt.shape[-2:]:
            pred = F.interpolate(pred, size=gt.shape[-2:], mode='bicubic', align_corners=False)
        loss = loss + F.mse_loss(pred, gt.cuda())
    return loss


@pytest.mark.parametrize('plain', [True, False])
def test_training_iteration_vs_oracle(plain, plain_ckpt):
    """One full iteration of the reference's training loop through the drop-in module in train() mode: loss and the
    gradient of every unique parameter after BPTT over the sequence, then two optimiser steps with the fused Adam
    against the oracle's restated Adam on its own fp32 gradients."""
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad
    b, h, w, steps = 2, 22, 40, 4
    gt_hw = (88, 160) if plain else (90, 160)                    # (90,160): the bicubic-resize branch of train.py:224-228
    sd = plain_ckpt if plain else O.surrogate_state_dict(plain=False, seed=2024, transplant=plain_ckpt)
    sd = {k: v.clone() for k, v in sd.items()}
    m = (BMCNet_plain if plain else BMCNet)(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    xs = [synth_counts(b, h, w, 700 + s) for s in range(steps)]
    g = torch.Generator().manual_seed(77)
    gts = [torch.poisson(torch.full((b, 2) + gt_hw, 0.3), generator=g) for _ in range(steps)]
    opt = FusedAdamAMSGrad(m.parameters())
    names = {O._alias_root(n): p for n, p in m.named_parameters()}
    opt.zero_grad()
    loss = _sequence(m, xs, gts, 1 if plain else 3)
    loss.backward()
    ref_loss, ref_grads, leaves = T.loss_and_grads(sd, xs, gts, plain)
    assert abs(loss.item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    report = {}
    for k, p in names.items():
        gr, rr = p.grad.detach().cpu(), ref_grads[k]
        assert torch.isfinite(gr).all(), k
        report[k] = (_rel_max(gr, rr), _rel_l2(gr, rr), rr.abs().max().item())
    worst = sorted(report.items(), key=lambda kv: -kv[1][0])[:5]
    try:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        os.makedirs(d, exist_ok=True)
        json.dump({'loss': loss.item(), 'ref_loss': ref_loss.item(), 'grads': report},
                  open(os.path.join(d, 'train_parity_%s.json' % ('plain' if plain else 'bmcnet')), 'w'))
    except OSError:
        pass
    for k, (emax, el2, mag) in report.items():
        assert emax <= 1.5e-2, (k, emax, el2, mag, worst)
        assert el2 <= 1e-2, (k, emax, el2, mag, worst)
    # optimiser: the fused kernel on OUR gradients vs the restated rule on the SAME gradients (isolates the kernel)
    before = {k: p.detach().cpu().clone() for k, p in names.items()}
    mine_grads = {k: p.grad.detach().cpu().clone() for k, p in names.items()}
    opt.step()
    ref_params = {k: v.clone() for k, v in before.items()}
    T.adam_amsgrad_step(ref_params, mine_grads, {})
    for k, p in names.items():
        assert torch.allclose(p.detach().cpu(), ref_params[k], rtol=1e-6, atol=2e-7), k
    # the next forward must see the updated weights (version bump -> fresh weight packs)
    opt.zero_grad()
    loss2 = _sequence(m, xs, gts, 1 if plain else 3)
    assert torch.isfinite(loss2) and loss2.item() != loss.item()
    loss2.backward()
    opt.step()


@pytest.mark.parametrize('plain', [True, False])
def test_graphed_iteration_matches_eager(plain, plain_ckpt):
    """GraphedIteration (zero_grad + sequence + backward replayed as one CUDA graph, then the fused Adam) against the
    same iteration issued op by op: identical loss and parameters over three iterations with fresh inputs each time --
    the weight packs inside the graph follow the optimiser's updates and the static input buffers are refilled."""
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration
    b, h, w, steps, iters = 2, 22, 40, 3, 3
    gt_hw = (88, 160) if plain else (90, 160)
    sd = plain_ckpt if plain else O.surrogate_state_dict(plain=False, seed=2024, transplant=plain_ckpt)
    g = torch.Generator().manual_seed(78)
    data = [([synth_counts(b, h, w, 900 + 10 * i + s).cuda() for s in range(steps)],
             [torch.poisson(torch.full((b, 2) + gt_hw, 0.3), generator=g).cuda() for _ in range(steps)]) for i in range(iters)]

    def fresh():
        m = (BMCNet_plain if plain else BMCNet)(4, 128, 5)
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        m = m.cuda().train()
        return m, FusedAdamAMSGrad(m.parameters(), lr=1e-3 if plain else 1e-4)

    m_e, opt_e = fresh()
    eager = []
    for xs, gts in data:
        opt_e.zero_grad()
        loss = _sequence(m_e, xs, gts, 1 if plain else 3)
        loss.backward()
        opt_e.step()
        eager.append(loss.item())
    m_g, opt_g = fresh()
    it = GraphedIteration(m_g, opt_g, *data[0])
    graphed = [it(xs, gts).item() for xs, gts in data]
    assert opt_g.step_count == iters
    for le, lg in zip(eager, graphed):
        assert abs(le - lg) <= 1e-5 * abs(le), (eager, graphed)
    assert eager[0] != eager[1] and all(np.isfinite(eager))
    # (cuBLAS picks other bmm algorithms under capture, so the two runs agree to rounding, not bit for bit; Adam turns a
    # rounding-level change of a near-zero gradient into a fraction of one lr-sized step)
    lr = opt_g.lr
    for (n, pe), pg in zip(m_e.named_parameters(), m_g.parameters()):
        d = (pe.detach() - pg.detach()).abs()
        # an element whose gradient is rounding noise can take a different +-lr step in each run: bound the bulk, not it
        assert d.mean().item() <= 0.01 * lr and d.max().item() <= 2 * lr * iters, (n, d.max().item(), d.mean().item())
        assert (d > 0.1 * lr).float().mean().item() <= 2e-2, (n, (d > 0.1 * lr).float().mean().item())
    moved = max((p.detach().cpu() - sd[n]).abs().max().item() for n, p in m_g.named_parameters() if n in sd)
    assert moved >= 2 * lr            # three Adam steps did move the weights


def test_graphed_iteration_side_stream_is_exact_and_reproducible(plain_ckpt):
    """The weight-gradient kernels of a captured iteration run on a side stream.  Gradient tensors they read must not be
    updated in place by the autograd engine meanwhile (models/_train.py, _ConvFn.backward): with and without the side
    stream, and from run to run, the flat gradient buffer is bit-identical over three iterations."""
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration
    b, h, w, steps = 2, 22, 40, 3
    sd = O.surrogate_state_dict(plain=False, seed=2024, transplant=plain_ckpt)
    xs = [synth_counts(b, h, w, 900 + s).cuda() for s in range(steps)]
    g = torch.Generator().manual_seed(3)
    gts = [torch.poisson(torch.full((b, 2, 4 * h, 4 * w), 0.3), generator=g).cuda() for _ in range(steps)]

    def run(side):
        m = BMCNet(4, 128, 5)
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        m = m.cuda().train()
        opt = FusedAdamAMSGrad(m.parameters())
        it = GraphedIteration(m, opt, xs, gts, wgrad_stream=side)
        out = []
        for _ in range(3):
            loss = it().item()
            out.append((loss, opt.grad.clone()))
        return out

    a, a2, c = run(True), run(True), run(False)
    for (la, ga), (la2, ga2), (lc, gc) in zip(a, a2, c):
        assert la == la2 == lc
        assert torch.equal(ga, ga2) and torch.equal(ga, gc)
    assert a[0][0] != a[1][0]


def test_graphed_iteration_overflow_backoff(plain_ckpt):
    """A loss scale far too large for fp16 (2^40) makes the very first gradient buffer non-finite: GraphedIteration must
    skip the optimiser step (parameters unchanged, no NaN), halve the scale and re-capture until the gradients are finite,
    then train normally."""
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration
    if _lib_act() != 'float16':
        pytest.skip('the loss scale only matters for fp16 activations')
    b, h, w, steps = 2, 22, 40, 2
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict({k: v.clone() for k, v in plain_ckpt.items()}, strict=True)
    m = m.cuda().train()
    m.loss_scale = 2.0 ** 40
    opt = FusedAdamAMSGrad(m.parameters())
    xs = [synth_counts(b, h, w, 950 + s).cuda() for s in range(steps)]
    g = torch.Generator().manual_seed(9)
    gts = [torch.poisson(torch.full((b, 2, 4 * h, 4 * w), 0.3), generator=g).cuda() for _ in range(steps)]
    before = opt.flat.clone()
    it = GraphedIteration(m, opt, xs, gts, warmup=1)
    it()
    assert it.overflows == 1 and opt.step_count == 0 and m.loss_scale == 2.0 ** 39
    assert torch.equal(opt.flat, before)                       # the step was skipped
    for _ in range(40):
        it()
        if opt.step_count > 0:
            break
    assert opt.step_count == 1 and it.overflows >= 2 and m.loss_scale < 2.0 ** 39
    assert torch.isfinite(opt.flat).all() and not torch.equal(opt.flat, before)
    l1 = it().item()
    assert np.isfinite(l1) and opt.step_count == 2
