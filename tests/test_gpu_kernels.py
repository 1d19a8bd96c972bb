"""GPU parity of the individual sm_100a kernels through the per-kernel C-ABI entry points
(bmc_conv_gemm, bmc_attention_weights, bmc_layernorm_rows, bmc_pack/unpack_nchw), against plain
PyTorch fp32 on the CPU with operands rounded to the kernels' 16-bit input type, and of the
sub-module forwards (ResidualBlock_noBN, LayerNorm2d, BIE) against the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def K():
    from bmcnet_esr_b200 import kernels
    return kernels


def _r(*shape, scale=1.0, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _q(t, K):
    from bmcnet_esr_b200 import _lib
    return t.to(_lib.act_dtype()).float()


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('cin,n,taps,relu,res,ln,b,h,w', [
    ([128], 128, 9, True, True, False, 2, 13, 21),       # ResidualBlock conv (submodules.py:31-35)
    ([128], 128, 1, False, False, False, 1, 7, 9),       # clustering / v 1x1 (submodules.py:63-67)
    ([128, 128], 128, 1, False, False, True, 2, 13, 21),  # convf + LayerNorm (submodules.py:63)
    ([128, 128], 128, 1, False, True, False, 2, 9, 30),  # unclustering + x_s (submodules.py:75)
    ([128, 64], 128, 9, True, False, False, 2, 13, 21),  # head conv with the 64-ch input tensor
    ([128, 128, 128], 128, 9, True, False, False, 1, 11, 13),
    ([128, 128], 32, 9, False, False, False, 2, 13, 21),  # conv_o (BMCNet.py:53)
    ([128], 128, 9, True, True, False, 3, 45, 80),       # NFS LR size, 93 tiles
    ([128], 128, 9, False, False, False, 2, 31, 56),     # EventZoom LR size
])
def test_conv_gemm(K, impl, cin, n, taps, relu, res, ln, b, h, w):
    k = 3 if taps == 9 else 1
    xs = [_q(_r(b, c, h, w, seed=10 + i), K) for i, c in enumerate(cin)]
    wt = _q(_r(n, sum(cin), k, k, scale=1.0 / (sum(cin) * taps) ** 0.5, seed=3), K)
    bias = _r(n, scale=0.1, seed=4)
    r = _q(_r(b, n, h, w, seed=5), K) if res else None
    y = F.conv2d(torch.cat(xs, 1), wt, bias, padding=k // 2)
    lnp = None
    if ln:
        gam, bet = 1 + _r(n, scale=0.1, seed=6), _r(n, scale=0.1, seed=7)
        mu = y.mean(1, keepdim=True)
        y = (y - mu) / ((y - mu).pow(2).mean(1, keepdim=True) + 1e-6).sqrt() * gam.view(1, -1, 1, 1) + bet.view(1, -1, 1, 1)
        lnp = (gam.cuda(), bet.cuda(), 1e-6)
    y = F.relu(y) if relu else y
    y = y + r if res else y
    segs, first = [], 0
    for c in cin:
        segs.append((first, c))
        first += c
    out, outf = K.conv_gemm([K.pack_nchw(x.cuda()) for x in xs], K.pack_conv_weight(wt.cuda(), segs), bias.cuda(),
                            b, h, w, taps, n=n, relu=relu, residual=K.pack_nchw(r.cuda()) if res else None,
                            ln=lnp, impl=impl, out_f32=True)
    R = K.rows_per_image(h, w)
    grid = outf.view(b, R, n)[:, :(h + 2) * (w + 2)].reshape(b, h + 2, w + 2, n)
    got32 = grid[:, 1:-1, 1:-1].permute(0, 3, 1, 2).cpu()
    # fp32 accumulation of exactly representable products: only summation order differs
    assert (got32 - y).abs().max().item() <= 2e-5 * max(1.0, y.abs().max().item())
    got16 = K.unpack_nchw(out, b, n, h, w).cpu()
    assert torch.equal(got16, _q(got32, K))                      # the 16-bit copy is the rounded fp32 result
    # halo and tail rows are forced to zero (they are the zero padding of the next conv)
    full = outf.view(b, R, n).clone()
    full[:, :(h + 2) * (w + 2)].view(b, h + 2, w + 2, n)[:, 1:-1, 1:-1] = 0
    assert float(full.abs().max()) == 0.0
    assert float(out.view(b, R, n)[:, (h + 2) * (w + 2):].float().abs().max() if R > (h + 2) * (w + 2) else 0.0) == 0.0


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('b,h,w,n_split', [(2, 13, 21, 3), (1, 45, 80, 8), (3, 31, 56, 1), (1, 5, 6, 4)])
def test_attention_weights_and_product(K, impl, b, h, w, n_split):
    """att = centres . v^T * nf^-0.5, softmax, out = softmax(att) . v  (submodules.py:69-73)."""
    c = _q(_r(b, 128, h, w, scale=0.5, seed=1), K)
    v = _q(_r(b, 128, h, w, scale=0.5, seed=2), K)
    scale = 128 ** -0.5
    att = torch.bmm(c.view(b, 128, -1), v.view(b, 128, -1).transpose(1, 2)) * scale
    pr = torch.softmax(att, -1)
    cp, vp = K.pack_nchw(c.cuda()), K.pack_nchw(v.cuda())
    probs, partial = K.attention_weights(cp, vp, b, h, w, scale, n_split, impl)
    got_att = partial.sum(1).cpu()
    assert (got_att - att).abs().max().item() <= 1e-5 * max(1.0, att.abs().max().item())
    got_p = probs.view(b, 2, 128, 64).permute(0, 2, 1, 3).reshape(b, 128, 128).float().cpu()
    assert (got_p - pr).abs().max().item() <= 2e-3            # 16-bit rounding of probabilities <= 1
    out = K.apply_dynamic_weights(vp, probs, b, h, w, impl=impl)
    ref = torch.bmm(got_p, v.view(b, 128, -1)).view(b, 128, h, w)
    got = K.unpack_nchw(out, b, 128, h, w).cpu()
    assert (got - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())


def test_pack_unpack_roundtrip_and_halo(K):
    x = _r(2, 128, 9, 14, seed=9)
    a = K.pack_nchw(x.cuda())
    assert torch.equal(K.unpack_nchw(a, 2, 128, 9, 14).cpu(), _q(x, K))
    R = K.rows_per_image(9, 14)
    g = a.view(2, R, 128)[:, :11 * 16].reshape(2, 11, 16, 128).float()
    assert float(g[:, 0].abs().max()) == 0 and float(g[:, -1].abs().max()) == 0
    assert float(g[:, :, 0].abs().max()) == 0 and float(g[:, :, -1].abs().max()) == 0


def _tol(K):
    from bmcnet_esr_b200 import _lib
    return 4e-3 if _lib.act_dtype() == torch.float16 else 3e-2


def test_residual_block_module_vs_oracle(K):
    from oracle import bmcnet_fp32 as O
    from bmcnet_esr_b200.models.submodules import ResidualBlock_noBN
    torch.manual_seed(0)
    m = ResidualBlock_noBN(128)
    m.conv1.weight.data.mul_(5)
    m.conv2.weight.data.mul_(5)
    x = _r(2, 128, 10, 12, seed=1)
    ref = O.resblock({'p.' + k: v for k, v in m.state_dict().items()}, 'p', x)
    got = m.cuda()(x.cuda()).cpu()
    assert (got - ref).abs().max().item() <= _tol(K) * ref.abs().max().item()


def test_layernorm2d_module_vs_oracle(K):
    from oracle import bmcnet_fp32 as O
    from bmcnet_esr_b200.models.submodules import LayerNorm2d
    m = LayerNorm2d(128)
    m.weight.data = 1 + _r(128, scale=0.1, seed=2)
    m.bias.data = _r(128, scale=0.1, seed=3)
    x = _r(2, 128, 6, 7, seed=4) * 3 + 1
    ref = O.layernorm2d(x, m.weight.data, m.bias.data, 1e-6)
    got = m.cuda()(x.cuda()).cpu()
    assert (got - ref).abs().max().item() <= 2 * _tol(K)


def test_bie_module_vs_oracle(K, plain_ckpt):
    """BIE with the trained weights of the shipped checkpoint (submodules.py:58-77)."""
    from oracle import bmcnet_fp32 as O
    from bmcnet_esr_b200.models.submodules import BIE
    m = BIE(128)
    pre = 'neuro.para_reschunk.0.'
    m.load_state_dict({k[len(pre):]: v for k, v in plain_ckpt.items() if k.startswith(pre)}, strict=True)
    x1, x2, xs = (F.relu(_r(2, 128, 12, 17, seed=s)) for s in (1, 2, 3))
    ref = O.bie(plain_ckpt, pre[:-1], x1, x2, xs)
    got = m.cuda()(x1.cuda(), x2.cuda(), xs.cuda())
    for g, r in zip(got, ref):
        assert (g.cpu() - r).abs().max().item() <= _tol(K) * r.abs().max().item()
