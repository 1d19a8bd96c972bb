"""Host logic of the training path (bmcnet_esr_b200/models/_train.py) WITHOUT a GPU: the kernels it calls
(bmc_conv_gemm, bmc_conv_wgrad, bmc_relu_backward, bmc_layernorm_rows and its backward, reached through kernels.py) are replaced by plain-torch fp32
stand-ins of their documented contracts (include/bmc_b200.h), and the resulting loss / gradients of a whole
BPTT sequence are compared with the fp32 autograd oracle (oracle/train_step.py, itself pinned to the reference's
modules + nn.MSELoss + torch.optim.Adam).  What this pins: the autograd wiring, the K-segment channel maps of every
`torch.cat` input, the mirrored / transposed data-gradient weights, the padded-NHWC layout conversions, the
loss-scale boundaries and the direct accumulation into the shared (aliased) Parameters' `.grad`.
The real kernels are checked against the same oracle in tests/test_gpu_train.py."""
import numpy as np
import pytest
import torch

from oracle import bmcnet_fp32 as O
from oracle import train_step as T
from oracle.make_golden import synth_counts


def _interior_mask(b, h, w):
    from bmcnet_esr_b200 import kernels as K
    r = K.rows_per_image(h, w)
    m = torch.zeros(r, dtype=torch.bool)
    idx = torch.arange((h + 2) * (w + 2))
    py, px = idx // (w + 2), idx % (w + 2)
    m[:(h + 2) * (w + 2)] = (py >= 1) & (py <= h) & (px >= 1) & (px <= w)
    return m.repeat(b)


def _shift(x, off):
    """y[r] = x[r + off], zeros outside."""
    y = torch.zeros_like(x)
    n = x.shape[0]
    if off >= 0:
        y[:n - off] = x[off:]
    else:
        y[-off:] = x[:n + off]
    return y


def _tap_offsets(taps, w):
    return [(t // 3 - 1) * (w + 2) + (t % 3 - 1) for t in range(9)] if taps == 9 else [0]


def fake_conv_gemm(srcs, wpk, bias, b, h, w, taps, n=128, relu=False, residual=None, ln=None, impl=0, out_f32=False):
    assert n == 128 and residual is None and ln is None
    wk = wpk.permute(1, 0, 2).reshape(128, -1).float()                 # [N, K], K = (segment, tap, channel)
    offs = _tap_offsets(taps, w)
    out = torch.zeros(srcs[0].shape[0], 128)
    k0 = 0
    for s in srcs:
        c = s.shape[1]
        for t, off in enumerate(offs):
            out += _shift(s.float(), off) @ wk[:, k0 + t * c:k0 + (t + 1) * c].t()
        k0 += taps * c
    assert k0 == wk.shape[1]
    if bias is not None:
        out = out + bias.view(1, -1)
    if relu:
        out = out.clamp(min=0)
    out[~_interior_mask(b, h, w)] = 0                                   # halo / tail rows forced to zero
    return out.to(srcs[0].dtype)


def fake_conv_gemm_stack(srcs, slabs, wstack, biases, n_jobs, b, h, w, taps, relu=False):
    from bmcnet_esr_b200 import kernels as K
    rows = b * K.rows_per_image(h, w)
    outs = []
    for j in range(n_jobs):
        wj = wstack[:, j * 128:(j + 1) * 128]
        sj = [s[slabs[i][j] * rows:(slabs[i][j] + 1) * rows] for i, s in enumerate(srcs)]
        outs.append(fake_conv_gemm(sj, wj, None if biases is None else biases[j], b, h, w, taps, relu=relu))
    return torch.cat(outs, 0)


def fake_conv_wgrad(dy, x, taps, b, h, w, cmap, cin_total, n_out, scale, grad_w, grad_b, workspace, n_split):
    offs = _tap_offsets(taps, w)
    keep = cmap >= 0
    for t, off in enumerate(offs):
        g = dy.float().t() @ _shift(x.float(), off)                     # [128 co, x_ch]
        gw = grad_w.view(n_out, cin_total, taps)
        gw[:, cmap[keep].long(), t] += scale * g[:n_out][:, keep]
    if grad_b is not None:
        grad_b += scale * dy.float().sum(0)[:n_out]


def fake_relu_backward(dy, y):
    return dy * (y > 0)


def fake_layernorm_rows(a, gamma, beta, eps):
    x = a.float()
    mu = x.mean(1, keepdim=True)
    var = ((x - mu) ** 2).mean(1, keepdim=True)
    return (gamma.view(1, -1) * ((x - mu) / (var + eps).sqrt()) + beta.view(1, -1)).to(a.dtype)


def fake_layernorm_rows_backward(x, dy, gamma, eps, scale, grad_gamma, grad_beta, workspace):
    """submodules.py:142-154"""
    xf, d = x.float(), dy.float()
    mu = xf.mean(1, keepdim=True)
    var = ((xf - mu) ** 2).mean(1, keepdim=True)
    rstd = 1.0 / (var + eps).sqrt()
    yh = (xf - mu) * rstd
    g = d * gamma.view(1, -1)
    dx = rstd * (g - yh * (g * yh).mean(1, keepdim=True) - g.mean(1, keepdim=True))
    grad_gamma += scale * (d * yh).sum(0)
    grad_beta += scale * d.sum(0)
    return dx.to(dy.dtype)


@pytest.fixture
def fake_kernels(monkeypatch):
    from bmcnet_esr_b200 import _lib, kernels as K
    monkeypatch.setattr(_lib, 'act_dtype', lambda: torch.float32)
    monkeypatch.setattr(K, 'conv_gemm', fake_conv_gemm)
    monkeypatch.setattr(K, 'conv_gemm_stack', fake_conv_gemm_stack)
    monkeypatch.setattr(K, 'conv_wgrad', fake_conv_wgrad)
    monkeypatch.setattr(K, 'relu_backward', fake_relu_backward)
    monkeypatch.setattr(K, 'layernorm_rows', fake_layernorm_rows)
    monkeypatch.setattr(K, 'layernorm_rows_backward', fake_layernorm_rows_backward)


def _grads_of(model):
    names = {}
    for n, p in model.named_parameters():
        names[O._alias_root(n)] = p
    return names


@pytest.mark.parametrize('stacked', [True, False])
@pytest.mark.parametrize('plain', [True, False])
def test_training_graph_matches_autograd_oracle(fake_kernels, plain, stacked, monkeypatch):
    # both launch structures: independent convolutions stacked into multi-job launches (small batches) / one launch each
    from bmcnet_esr_b200.models import _train as TR
    monkeypatch.setattr(TR, 'STACKED', stacked)
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    b, h, w, steps = 2, 6, 9, 2
    sd = O.surrogate_state_dict(plain=plain, seed=31)
    m = (BMCNet_plain if plain else BMCNet)(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m.train()
    m.loss_scale = 256.0                                                 # exercised even though the stand-ins are fp32
    xs = [synth_counts(b, h, w, 40 + s) for s in range(steps)]
    g = torch.Generator().manual_seed(5)
    gts = [torch.poisson(torch.full((b, 2, 4 * h, 4 * w), 0.3), generator=g) for _ in range(steps)]
    # ---- product graph (reference loop, train.py:206-236)
    n_state = 1 if plain else 3
    st = [torch.zeros(b, 128, h, w) for _ in range(n_state)] + [torch.zeros(b, 32, h, w)]
    loss, init = 0, True
    for x, gt in zip(xs, gts):
        st = list(m(x, *st, init))
        init = False
        loss = loss + torch.nn.functional.mse_loss(st[-1], gt)
    loss.backward()
    # ---- oracle
    ref_loss, ref_grads, _ = T.loss_and_grads(sd, xs, gts, plain)
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    got = _grads_of(m)
    assert sorted(got) == sorted(ref_grads)
    for k, p in got.items():
        assert p.grad is not None, k
        r = ref_grads[k]
        err = (p.grad - r).abs().max().item()
        assert err <= 2e-4 * r.abs().max().item() + 1e-9, (k, err, r.abs().max().item())
