"""Building blocks of BMCNet with the reference's module tree (reference: models/submodules.py).

Each class owns the same parameters under the same names as its reference counterpart, so
`state_dict()` / `load_state_dict(strict=True)` are interchangeable with the reference
(SURVEY.md F4, section 8b).  The `forward` of a block runs the sm_100a kernels through the
per-kernel C-ABI entry points (bmc_conv_gemm, bmc_attention_weights); the full models do not
call these forwards -- they hand the whole step to bmc_model_forward (see _engine.py).
Inference only: no autograd graph is recorded.
"""
import torch
import torch.nn as nn
from torch.nn import init

from .. import kernels as K


def _conv(nf_in, nf_out, k):
    return nn.Conv2d(nf_in, nf_out, k, 1, k // 2, bias=True)


def initialize_weights(net_l, scale=0.1):
    """Kaiming-normal(fan_in) * scale, zero bias (reference submodules.py:107-124)."""
    for net in (net_l if isinstance(net_l, list) else [net_l]):
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                init.kaiming_normal_(m.weight, a=0, mode='fan_in')
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias.data, 0.0)


def _run_conv(conv, xs, relu=False, residual=None, ln=None):
    """conv(cat(xs, 1)) on packed bf16 sources via one conv-gemm job."""
    b, h, w = xs[0][1:]
    taps = conv.kernel_size[0] * conv.kernel_size[1]
    segs, first = [], 0
    for a, _, _, _ in xs:
        segs.append((first, a.shape[1]))
        first += a.shape[1]
    wpk = K.pack_conv_weight(conv.weight.detach().float(), segs)
    return K.conv_gemm([a for a, _, _, _ in xs], wpk, conv.bias.detach(), b, h, w, taps, n=conv.out_channels,
                       relu=relu, residual=residual, ln=ln)


class ResidualBlock_noBN(nn.Module):
    """x + conv2(relu(conv1(x)))  (reference submodules.py:17-35)."""

    def __init__(self, nf=64):
        super().__init__()
        self.conv1 = _conv(nf, nf, 3)
        self.conv2 = _conv(nf, nf, 3)
        initialize_weights([self.conv1, self.conv2], 0.1)

    def _packed(self, a, b, h, w):
        t = _run_conv(self.conv1, [(a, b, h, w)], relu=True)
        return _run_conv(self.conv2, [(t, b, h, w)], residual=a)

    @torch.no_grad()
    def forward(self, x):
        b, c, h, w = x.shape
        return K.unpack_nchw(self._packed(K.pack_nchw(x), b, h, w), b, c, h, w)


class LayerNorm2d(nn.Module):
    """Per-pixel LayerNorm over channels, eps inside the sqrt (reference submodules.py:127-166)."""

    def __init__(self, channels, eps=1e-6):
        super().__init__()
        self.register_parameter('weight', nn.Parameter(torch.ones(channels)))
        self.register_parameter('bias', nn.Parameter(torch.zeros(channels)))
        self.eps = eps

    @torch.no_grad()
    def forward(self, x):
        b, c, h, w = x.shape
        if c != 128:
            raise NotImplementedError('LayerNorm2d kernel is specialised for 128 channels')
        y = K.layernorm_rows(K.pack_nchw(x), self.weight.detach(), self.bias.detach(), self.eps)
        # halo rows hold `bias` after the norm; unpack only reads interior pixels
        return K.unpack_nchw(y, b, c, h, w)


class BIE(nn.Module):
    """Bilateral information exchange (reference submodules.py:38-77)."""

    def __init__(self, nf=64):
        super().__init__()
        self.conv1 = ResidualBlock_noBN(nf)
        self.conv2 = self.conv1                       # shared weights, aliased state_dict keys
        self.convf1 = _conv(nf * 2, nf, 1)
        self.convf2 = self.convf1
        self.scale = nf ** -0.5
        self.norm_s = LayerNorm2d(nf)
        self.clustering = _conv(nf, nf, 1)
        self.unclustering = _conv(nf * 2, nf, 1)
        self.v1 = _conv(nf, nf, 1)
        self.v2 = _conv(nf, nf, 1)
        initialize_weights([self.convf1, self.convf2, self.clustering, self.unclustering, self.v1, self.v2], 0.1)

    @torch.no_grad()
    def forward(self, x_1, x_2, x_s):
        b, c, h, w = x_1.shape
        if c != 128:
            raise NotImplementedError('BIE kernels are specialised for nf=128')
        g = (b, h, w)
        a1, a2, a_s = K.pack_nchw(x_1), K.pack_nchw(x_2), K.pack_nchw(x_s)
        r1 = self.conv1._packed(a1, *g)
        r2 = self.conv2._packed(a2, *g)
        ln = (self.norm_s.weight.detach(), self.norm_s.bias.detach(), self.norm_s.eps)
        c1 = _run_conv(self.clustering, [(_run_conv(self.convf1, [(a_s, *g), (a2, *g)], ln=ln), *g)])
        c2 = _run_conv(self.clustering, [(_run_conv(self.convf2, [(a_s, *g), (a1, *g)], ln=ln), *g)])
        v1 = _run_conv(self.v1, [(a1, *g)])
        v2 = _run_conv(self.v2, [(a2, *g)])
        p1, _ = K.attention_weights(c1, v1, *g, self.scale)
        p2, _ = K.attention_weights(c2, v2, *g, self.scale)
        o1 = K.apply_dynamic_weights(v1, p1, *g, residual=r2)
        o2 = K.apply_dynamic_weights(v2, p2, *g, residual=r1)
        s = _run_conv(self.unclustering, [(c1, *g), (c2, *g)], residual=a_s)
        return tuple(K.unpack_nchw(t, b, c, h, w) for t in (o1, o2, s))


def pixel_unshuffle(input, upscale_factor):
    """Inverse of pixel_shuffle: channel = c*r*r + ry*r + rx (reference submodules.py:80-92)."""
    return torch.nn.functional.pixel_unshuffle(input, upscale_factor)


class PixelUnShuffle(nn.Module):
    def __init__(self, upscale_factor):
        super().__init__()
        self.upscale_factor = upscale_factor

    def forward(self, input):
        return pixel_unshuffle(input, self.upscale_factor)

    def extra_repr(self):
        return 'upscale_factor={}'.format(self.upscale_factor)
