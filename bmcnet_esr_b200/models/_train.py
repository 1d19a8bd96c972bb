"""Training path of BMCNet / BMCNet_plain (reference train.py:202-237).

The reference trains with plain autograd over its nn.Modules: forward over a sequence of window pairs with the
state carried (no detach -> BPTT), `loss += MSELoss(pred, gt)`, one `loss.backward()`, one Adam(amsgrad) step.
Here the SAME user code works (`model.train()`, the loop, `loss.backward()`, any torch optimiser): when autograd
is recording, `BMCNet.forward` / `BMCNet_plain.forward` route to this module, which builds the graph out of
`torch.autograd.Function`s whose forward AND backward are the sm_100a kernels of libbmc_b200:

  every convolution (216 per BMCNet step, 99.5 % of the FLOPs; BMCNet.py:40-53, submodules.py:25-26,44-53)
      forward   bmc_conv_gemm (the tcgen05 slab / per-tap kernels, K segments = torch.cat inputs, bias + ReLU fused)
      dgrad     bmc_conv_gemm again: dY convolved with the spatially mirrored, channel-transposed weights
      wgrad     bmc_conv_wgrad (csrc/train.cu: tcgen05 split-K over pixels, fp32, accumulated straight into
                `param.grad` -- aliased modules share one Parameter, hence one gradient buffer)
      ReLU'     bmc_relu_backward
  channel LayerNorm (norm_s of every BIE, submodules.py:127-166): bmc_layernorm_rows / bmc_layernorm_rows_backward
  the small rest (the 128x128 attention bmm/softmax, residual adds, pixel (un)shuffle, bilinear base, MSE) is ordinary
  differentiable PyTorch on the padded-NHWC tensors, in fp32 where it matters.

Activations and activation gradients are act16 (fp16 in the default build); weight gradients accumulate in fp32.
fp16 gradients of a mean-reduced MSE over 10^5 outputs (~1e-6) would underflow, so the whole 16-bit region runs under a
STATIC loss scale that never leaves this module: gradients are multiplied by `loss_scale` where they enter (the
prediction / state outputs) and divided where they leave (wgrad's `scale`, the LayerNorm parameters, the state
inputs), so `loss.backward()` on the user's unscaled loss yields unscaled `param.grad`.

Also here: FusedAdamAMSGrad (bmc_adam_amsgrad_step over one flat fp32 buffer) and allreduce_gradients (the one
collective of the data-parallel training step: a single NCCL all-reduce of the 2,731,680 alias-deduplicated
gradients).
"""
import ctypes as C

import torch
import torch.nn.functional as F

from .. import _lib, kernels as K
from .._lib import check, lib, stream_ptr

WGRAD_SPLITS = 16
STACKED = True       # independent convolutions of one shape as multi-job launches on stacked tensors (conv_stack) while one
                     # job fills at most half a wave of SMs (else one launch each: the gather copies cost more than they save)


def _use_stack(b, h, w):
    return STACKED and 2 * b * K.rows_per_image(h, w) <= 256 * 148
DEFAULT_LOSS_SCALE = 2.0 ** 14


# ------------------------------------------------------------------------------------------ layout (differentiable)
def to_packed(x, c_pad):
    """fp32 NCHW [B,C,H,W] -> padded NHWC act16 [B*R, c_pad] with zero halo, built from differentiable torch ops."""
    b, c, h, w = x.shape
    r = K.rows_per_image(h, w)
    y = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(b, (h + 2) * (w + 2), c)
    y = F.pad(y, (0, c_pad - c, 0, r - (h + 2) * (w + 2)))
    return y.reshape(b * r, c_pad).to(_lib.act_dtype())


def from_packed(a, b, c, h, w):
    """padded NHWC act16 [B*R, c_pad] -> fp32 NCHW [B,C,H,W] (differentiable)."""
    r = K.rows_per_image(h, w)
    y = a.view(b, r, a.shape[1])[:, :(h + 2) * (w + 2), :c].reshape(b, h + 2, w + 2, c)
    return y[:, 1:h + 1, 1:w + 1].permute(0, 3, 1, 2).float()


class _ScaleGrad(torch.autograd.Function):
    """identity forward; backward multiplies the gradient by a constant (loss-scale boundaries)."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.s, None


# ------------------------------------------------------------------------------------------ convolution
class _Ctx:
    """Per-model training context: geometry-independent caches (packed weights by parameter version, channel
    maps, wgrad workspace) and the loss scale."""

    def __init__(self, loss_scale):
        self.loss_scale = float(loss_scale)
        self.wcache = {}
        self.cmaps = {}
        self.gidx = {}
        self.ws = None
        self.ln_ws = None
        self.side = None          # GraphedIteration: the stream the weight-gradient kernels are recorded on
        self.keep = []            # gradient tensors the side stream reads (see _ConvFn.backward)
        self._anchor = None

    def workspace(self, dev):
        nbytes = lib().bmc_conv_wgrad_workspace_bytes(WGRAD_SPLITS, 9, 128)
        if self.ws is None or self.ws.device != dev:
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self.ws

    def new_anchor(self, dev):
        """A fresh leaf that requires grad, one per forward call (see _ConvFn)."""
        self._anchor = torch.zeros((), device=dev, requires_grad=True)
        return self._anchor

    @property
    def anchor(self):
        if self._anchor is None:            # building blocks used on their own (tests): any CUDA leaf will do
            self.new_anchor(torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else 'cpu')
        return self._anchor

    def ln_workspace(self, dev):
        if self.ln_ws is None or self.ln_ws.device != dev:
            self.ln_ws = torch.empty(lib().bmc_layernorm_rows_backward_workspace_bytes(), dtype=torch.uint8, device=dev)
        return self.ln_ws

    def cmap(self, idx, dev):
        key = (tuple(idx), str(dev))
        if key not in self.cmaps:
            self.cmaps[key] = torch.tensor(idx, dtype=torch.int32, device=dev)
        return self.cmaps[key]

    def lidx(self, idx, dev):
        key = ('l', tuple(idx), str(dev))
        if key not in self.cmaps:
            self.cmaps[key] = torch.tensor(idx, dtype=torch.int64, device=dev)
        return self.cmaps[key]

    def gather_index(self, idx, dev):
        """(clamped int64 index, fp32 0/1 mask [1,c,1,1]) of a channel list; cached, so that packing weights issues no
        host-to-device copy (a CUDA-graph capture of the iteration, GraphedIteration, must not contain one)."""
        key = (tuple(idx), str(dev))
        hit = self.gidx.get(key)
        if hit is None:
            it = torch.tensor(idx, device=dev)
            hit = (it.clamp(min=0), (it >= 0).view(1, -1, 1, 1).float())
            self.gidx[key] = hit
        return hit

    def packed_weight(self, weight, segs, mode, seg=None):
        """mode 'fwd': chunk-major forward weights for the K segments `segs` (lists of weight input channels, -1 =
        padding), output channels zero-padded to 128.  mode 'bwd': the data-gradient weights of segment `seg`:
        out = its (padded) input channels, in = the 128 (padded) output channels, taps mirrored."""
        key = (id(weight), weight._version, tuple(map(tuple, segs)), mode, seg)
        hit = self.wcache.get(key)
        if hit is not None:
            return hit
        w = weight.detach().float()
        n, cin, kh, kw = w.shape
        if n < 128:
            w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, 128 - n))
        if mode == 'fwd':
            cols = []
            for idx in segs:
                it, mask = self.gather_index(idx, w.device)
                ws = w.index_select(1, it) * mask
                cols.append(ws.reshape(128, len(idx), kh * kw).permute(0, 2, 1).reshape(128, -1))     # [N, taps*c]
            wk = torch.cat(cols, 1)
        else:
            idx = segs[seg]
            it, mask = self.gather_index(idx, w.device)
            ws = w.index_select(1, it) * mask                                                         # [128 co, c, kh, kw]
            wt = ws.permute(1, 0, 2, 3).flip(2, 3)                                                    # [c, 128 co, kh, kw] mirrored
            c = len(idx)
            if c < 128:
                wt = F.pad(wt, (0, 0, 0, 0, 0, 0, 0, 128 - c))
            wk = wt.reshape(128, 128, kh * kw).permute(0, 2, 1).reshape(128, -1)
        k = wk.shape[1]
        out = wk.reshape(128, k // 64, 64).permute(1, 0, 2).contiguous().to(_lib.act_dtype())
        self.wcache[key] = out
        return out


def _packed_stack(tc, convs, segs, mode, seg=None):
    """The packs of J convolutions of one shape side by side: act16 [K/64, J*128, 64] (K.conv_gemm_stack)."""
    key = ('stack', tuple((id(c.weight), c.weight._version) for c in convs), tuple(map(tuple, segs)), mode, seg)
    hit = tc.wcache.get(key)
    if hit is None:
        hit = torch.cat([tc.packed_weight(c.weight, segs, mode, seg) for c in convs], 1).contiguous()
        tc.wcache[key] = hit
    return hit


def _grad_buf(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
    return p.grad


class _ConvFn(torch.autograd.Function):
    """`weight` is a tensor input only so that the output requires grad (the very first convolutions of a sequence have
    no other input that does): the caller passes the forward call's ANCHOR, a fresh 0-d leaf (`_Ctx.new_anchor`), not
    the Parameter.  The parameter gradients -- weight and bias -- are accumulated into `.grad` by the wgrad kernel and
    backward returns None for the anchor, so no Parameter ever enters the autograd graph and no AccumulateGrad node of a
    Parameter (which remembers the stream it was created on, and outlives an iteration whenever a caller keeps a loss
    tensor) can tie a captured iteration to another stream."""

    @staticmethod
    def forward(ctx, weight, tc, conv, segs, relu, use_bias, geom, *srcs):
        b, h, w = geom
        taps = conv.kernel_size[0] * conv.kernel_size[1]
        n_out = conv.out_channels
        wpk = tc.packed_weight(conv.weight, segs, 'fwd')
        bias = None
        if use_bias:
            bias = conv.bias.detach().float()
            if n_out < 128:
                bias = F.pad(bias, (0, 128 - n_out))
        srcs = [s.contiguous() for s in srcs]
        out = K.conv_gemm(srcs, wpk, bias, b, h, w, taps, n=128, relu=relu)
        ctx.tc, ctx.conv, ctx.segs, ctx.relu, ctx.geom, ctx.taps, ctx.use_bias = tc, conv, segs, relu, geom, taps, use_bias
        ctx.save_for_backward(out if relu else None, *srcs)
        return out

    @staticmethod
    def backward(ctx, dout):
        tc, conv, segs, geom, taps = ctx.tc, ctx.conv, ctx.segs, ctx.geom, ctx.taps
        b, h, w = geom
        out, *srcs = ctx.saved_tensors
        dz = dout.contiguous()
        dev = dz.device
        if ctx.relu:
            dz = K.relu_backward(dz, out)
        ws = tc.workspace(dev)
        gw = _grad_buf(conv.weight)
        gb = _grad_buf(conv.bias)
        inv = 1.0 / tc.loss_scale

        def wgrads(dy):
            for i, (src, idx) in enumerate(zip(srcs, segs)):
                K.conv_wgrad(dy, src, taps, b, h, w, tc.cmap(idx, dev), conv.in_channels, conv.out_channels, inv,
                             gw, gb if (i == 0 and ctx.use_bias) else None, ws, WGRAD_SPLITS)

        if tc.side is None:
            wgrads(dz)
        else:
            # Nothing downstream in the backward pass reads a weight gradient, so inside a captured iteration the wgrad
            # kernels go to ONE side stream (they stay ordered among themselves: shared workspace, aliased gradient
            # buffers) and overlap the dgrad chain, which at training batches fills a fifth of the SMs.
            # Without a ReLU `dz` IS the incoming gradient tensor (the backward of `x + conv(t)` hands the same tensor to
            # both addends), and the autograd engine accumulates further gradients into such a buffer IN PLACE on the main
            # stream once it holds the last reference (input_buffer.cpp: use_count() == 1) -- while the side stream may
            # still be reading it.  Holding a reference until the iteration's backward is over keeps the engine on the
            # out-of-place path for this tensor (same number of kernels, no copy).
            wg_dz = dz
            if not ctx.relu:
                tc.keep.append(dz)
            tc.side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(tc.side):
                wgrads(wg_dz)
            for t in (wg_dz, *srcs):
                t.record_stream(tc.side)
        grads = []
        for i, src in enumerate(srcs):
            if not ctx.needs_input_grad[7 + i]:
                grads.append(None)
                continue
            wt = tc.packed_weight(conv.weight, segs, 'bwd', i)
            g = K.conv_gemm([dz], wt, None, b, h, w, taps, n=128)
            grads.append(g if src.shape[1] == 128 else g[:, :src.shape[1]].contiguous())
        return (None, None, None, None, None, None, None, *grads)


def conv(tc, module, srcs, segs, geom, relu=False):
    """module(cat(srcs)) on packed act16 sources; segs[i] lists, per channel of source i, the input channel of
    `module.weight` it multiplies (-1 for padding channels).  bmc_conv_gemm takes up to 3 sources per job: a wider
    concatenation (conv_fs of BMCNet: 4) is the sum of two launches, bias in the first, ReLU after the sum."""
    if len(srcs) <= 3:
        return _ConvFn.apply(tc.anchor, tc, module, segs, relu, True, geom, *srcs)
    a = _ConvFn.apply(tc.anchor, tc, module, segs[:3], False, True, geom, *srcs[:3])
    c = _ConvFn.apply(tc.anchor, tc, module, segs[3:], False, False, geom, *srcs[3:])
    out = a + c
    return F.relu(out) if relu else out


class _ConvStackFn(torch.autograd.Function):
    """J convolutions of one shape (different weights, or one weight applied to J inputs) as ONE multi-job launch, forward
    and data-gradient alike, on stacked operands: source tensor s is [n_slabs_s * rows, C_s] and segment s of job j reads
    slab `slabs[s][j]`; the result is [J * rows, 128].  At training batches one job fills a fifth of the SMs and its
    launch costs the one-wave floor (14 us): four jobs per launch cost the same.  Weight gradients: per job, as _ConvFn."""

    @staticmethod
    def forward(ctx, tc, convs, segs, slabs, relu, geom, n_w, *tensors):
        b, h, w = geom
        srcs = [t.contiguous() for t in tensors[n_w:]]
        taps = convs[0].kernel_size[0] * convs[0].kernel_size[1]
        wst = _packed_stack(tc, convs, segs, 'fwd')
        biases = []
        for c in convs:
            bias = c.bias.detach().float()
            biases.append(F.pad(bias, (0, 128 - c.out_channels)) if c.out_channels < 128 else bias)
        out = K.conv_gemm_stack(srcs, slabs, wst, biases, len(convs), b, h, w, taps, relu=relu)
        ctx.tc, ctx.convs, ctx.segs, ctx.slabs, ctx.relu, ctx.geom, ctx.taps, ctx.n_w = tc, convs, segs, slabs, relu, geom, taps, n_w
        ctx.save_for_backward(out if relu else None, *srcs)
        return out

    @staticmethod
    def backward(ctx, dout):
        tc, convs, segs, slabs, geom, taps = ctx.tc, ctx.convs, ctx.segs, ctx.slabs, ctx.geom, ctx.taps
        b, h, w = geom
        out, *srcs = ctx.saved_tensors
        J = len(convs)
        rows = b * K.rows_per_image(h, w)
        dz = dout.contiguous()
        dev = dz.device
        if ctx.relu:
            dz = K.relu_backward(dz, out)
        ws = tc.workspace(dev)
        inv = 1.0 / tc.loss_scale

        def wgrads(dy):
            for j, c in enumerate(convs):
                gw, gb = _grad_buf(c.weight), _grad_buf(c.bias)
                dyj = dy[j * rows:(j + 1) * rows]
                for i, (src, idx) in enumerate(zip(srcs, segs)):
                    sl = slabs[i][j]
                    K.conv_wgrad(dyj, src[sl * rows:(sl + 1) * rows], taps, b, h, w, tc.cmap(idx, dev), c.in_channels,
                                 c.out_channels, inv, gw, gb if i == 0 else None, ws, WGRAD_SPLITS)

        if tc.side is None:
            wgrads(dz)
        else:                                   # see _ConvFn.backward
            if not ctx.relu:
                tc.keep.append(dz)
            tc.side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(tc.side):
                wgrads(dz)
            for t in (dz, *srcs):
                t.record_stream(tc.side)
        # Data gradients, one multi-job launch per segment, routed back to the slabs the jobs read.  Cheap routes first:
        # identity; a permutation of the slabs (one gather); every slab read by r consecutive jobs (one sum); else
        # zeros + index_add.  A tensor passed for several segments (unclustering reads c for both) collects all of them in
        # ONE buffer, returned for its first occurrence (None = zero for the others).
        grads = [None] * len(srcs)
        ident = list(range(J))
        first_of = {}
        for i, src in enumerate(srcs):
            if not ctx.needs_input_grad[7 + ctx.n_w + i]:
                continue
            wt = _packed_stack(tc, convs, segs, 'bwd', i)
            g = K.conv_gemm_stack([dz], [ident], wt, None, J, b, h, w, taps)             # [J * rows, 128]
            c = src.shape[1]
            if c != 128:
                g = g[:, :c]
            n_slabs = src.shape[0] // rows
            sl = list(slabs[i])
            key = (src.data_ptr(), src.shape[0], c)
            shared = sum(1 for k2, s2 in enumerate(srcs) if (s2.data_ptr(), s2.shape[0], s2.shape[1]) == key) > 1
            if shared:
                j0 = first_of.get(key)
                if j0 is None:
                    first_of[key] = i
                    grads[i] = torch.zeros(n_slabs, rows, c, dtype=g.dtype, device=dev)
                    j0 = i
                grads[j0].index_add_(0, tc.cmap(sl, dev), g.reshape(J, rows, c))
            elif sl == list(range(n_slabs)):
                grads[i] = g.contiguous()
            elif J == n_slabs and sorted(sl) == list(range(n_slabs)):                      # permutation: grad[sl[j]] = g[j]
                inv = [0] * J
                for j, t in enumerate(sl):
                    inv[t] = j
                grads[i] = g.reshape(J, rows, c).index_select(0, tc.lidx(inv, dev))
            elif J % n_slabs == 0 and sl == [j // (J // n_slabs) for j in range(J)]:       # slab t read by jobs t r .. t r + r - 1
                grads[i] = g.reshape(n_slabs, J // n_slabs, rows, c).sum(1)
            else:
                gs = torch.zeros(n_slabs, rows, c, dtype=g.dtype, device=dev)
                gs.index_add_(0, tc.cmap(sl, dev), g.reshape(J, rows, c))
                grads[i] = gs
        grads = [None if g is None else g.reshape(-1, g.shape[-1]) for g in grads]
        return (None, None, None, None, None, None, None) + (None,) * ctx.n_w + tuple(grads)


def conv_stack(tc, convs, srcs, segs, slabs, geom, relu=False):
    """convs[j](cat of the slabs slabs[s][j] of srcs[s]) for all j in one launch -> [J * rows, 128]."""
    ws = [tc.anchor]
    return _ConvStackFn.apply(tc, convs, segs, slabs, relu, geom, len(ws), *ws, *srcs)


def resblock_stack(tc, mods, x, geom):
    """ResidualBlock_noBN mods[j] on slab j of x."""
    ident = [list(range(len(mods)))]
    t = conv_stack(tc, [m.conv1 for m in mods], [x], [_R128], ident, geom, relu=True)
    return x + conv_stack(tc, [m.conv2 for m in mods], [t], [_R128], ident, geom)


def _take(tc, x, idx, rows):
    """slabs `idx` (a list) of a stacked tensor (differentiable gather copy)."""
    return x.view(-1, rows, x.shape[1]).index_select(0, tc.lidx(idx, x.device)).reshape(len(idx) * rows, x.shape[1])


def bie_stack(tc, m, x, xs, n_inst, geom):
    """BIE.forward (submodules.py:58-77) for n_inst independent instances sharing the module's weights.
    x: [2 n_inst * rows, 128] = (x1, x2) per instance; xs: [n_inst * rows, 128].  Returns (x1', x2') stacked like x, and xs'."""
    b, h, w = geom
    r = K.rows_per_image(h, w)
    rows = b * r
    J = 2 * n_inst
    two = [_R128, _seg(128)]
    ident = list(range(J))
    res = resblock_stack(tc, [m.conv1, m.conv2] * n_inst, x, geom)                       # (Res(x1), Res(x2)) per instance
    # centres_k = clustering(LN(convf_k([x_s, x_other])))
    u = conv_stack(tc, [m.convf1, m.convf2] * n_inst, [xs, x], two,
                   [[i // 2 for i in ident], [i ^ 1 for i in ident]], geom)
    c = conv_stack(tc, [m.clustering] * J, [layernorm_rows(tc, m.norm_s, u)], [_R128], [ident], geom)
    v = conv_stack(tc, [m.v1, m.v2] * n_inst, [x], [_R128], [ident], geom)
    o = _AttendFn.apply(c, v, J * b, r, float(m.scale))                                    # every (instance, k, image) on its own
    ns = conv_stack(tc, [m.unclustering] * n_inst, [c, c], two,
                    [[2 * i for i in range(n_inst)], [2 * i + 1 for i in range(n_inst)]], geom) + xs
    return o + _take(tc, res, [i ^ 1 for i in ident], rows), ns                                                   # (out_1 + Res(x_2), out_2 + Res(x_1))


_R128 = list(range(128))


def _seg(first, count=128, pad_to=128):
    return list(range(first, first + count)) + [-1] * (pad_to - count)


# ------------------------------------------------------------------------------------------ blocks
def resblock(tc, m, x, geom):
    """ResidualBlock_noBN (submodules.py:31-35)."""
    t = conv(tc, m.conv1, [x], [_R128], geom, relu=True)
    return x + conv(tc, m.conv2, [t], [_R128], geom)


class _LayerNormFn(torch.autograd.Function):
    """LayerNorm2d / LayerNormFunction (submodules.py:127-166) on packed rows: forward = bmc_layernorm_rows, backward =
    bmc_layernorm_rows_backward (dx; dgamma / dbeta accumulated into `.grad`, un-scaled, like the conv weights)."""

    @staticmethod
    def forward(ctx, x, weight, tc, norm):
        x = x.contiguous()
        ctx.tc, ctx.norm = tc, norm
        ctx.save_for_backward(x)
        return K.layernorm_rows(x, norm.weight.detach(), norm.bias.detach(), norm.eps)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        tc, norm = ctx.tc, ctx.norm
        dx = K.layernorm_rows_backward(x, dy.contiguous(), norm.weight.detach(), norm.eps, 1.0 / tc.loss_scale,
                                       _grad_buf(norm.weight), _grad_buf(norm.bias), tc.ln_workspace(x.device))
        return dx, None, None, None


class _tf32_matmul:
    """The attention products run on the tensor cores as TF32 x TF32 -> fp32 (cuBLAS).  Their big operands (centres, v,
    incoming gradients) hold 16-bit values, which TF32 represents exactly, so those products are exact; fp32 operands
    (softmax probabilities, logit gradients) are rounded to 11 bits like every other activation of this path.  The
    default fp32 path of torch.bmm is a SIMT sgemm: 76 us per call at batch 2, a fifth of the whole iteration."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class _AttendFn(torch.autograd.Function):
    """out = (softmax(centres^T v * scale) v^T)^T per image (BIE.forward, submodules.py:69-73) on packed rows
    c, v: act16 [B*R, 128] (halo rows are zero and contribute nothing).  Logits and probabilities stay fp32 (the logits
    reach +-33: 16-bit logits would cost 3 % in the softmax)."""

    @staticmethod
    def forward(ctx, c, v, b, r, scale):
        cf, vf = c.view(b, r, 128).float(), v.view(b, r, 128).float()
        with _tf32_matmul():
            att = torch.bmm(cf.transpose(1, 2), vf)                        # [b, c, c'] = centres . v^T (:69-70)
            p = torch.softmax(att * scale, -1)
            out = torch.bmm(vf, p.transpose(1, 2))                         # (P v)^T in row layout (:72-73)
        ctx.save_for_backward(c, v, p)
        ctx.geom, ctx.scale = (b, r), scale
        return out.reshape(b * r, 128).to(c.dtype)

    @staticmethod
    def backward(ctx, g):
        c, v, p = ctx.saved_tensors
        (b, r), scale = ctx.geom, ctx.scale
        cf, vf, gf = c.view(b, r, 128).float(), v.view(b, r, 128).float(), g.reshape(b, r, 128).float()
        with _tf32_matmul():
            dp = torch.bmm(gf.transpose(1, 2), vf)                         # out[r,c] = sum_c' v[r,c'] p[c,c']
            datt = p * (dp - (dp * p).sum(-1, keepdim=True)) * scale       # softmax backward, then the logit scale
            dv = torch.baddbmm(torch.bmm(gf, p), cf, datt)                 # through `out` and through the logits
            dc = torch.bmm(vf, datt.transpose(1, 2))
        return dc.reshape(b * r, 128).to(c.dtype), dv.reshape(b * r, 128).to(v.dtype), None, None, None


def layernorm_rows(tc, norm, y):
    return _LayerNormFn.apply(y, tc.anchor, tc, norm)


def bie(tc, m, x1, x2, xs, geom):
    """BIE.forward (submodules.py:58-77) on packed tensors."""
    b, h, w = geom
    r = K.rows_per_image(h, w)
    r1 = resblock(tc, m.conv1, x1, geom)
    r2 = resblock(tc, m.conv2, x2, geom)
    two = [_R128, _seg(128)]

    def centre(convf, other):
        u = conv(tc, convf, [xs, other], two, geom)
        return conv(tc, m.clustering, [layernorm_rows(tc, m.norm_s, u)], [_R128], geom)

    c1, c2 = centre(m.convf1, x2), centre(m.convf2, x1)
    v1 = conv(tc, m.v1, [x1], [_R128], geom)
    v2 = conv(tc, m.v2, [x2], [_R128], geom)

    def attend(c, v):
        return _AttendFn.apply(c, v, b, r, float(m.scale))

    o1, o2 = attend(c1, v1), attend(c2, v2)
    ns = conv(tc, m.unclustering, [c1, c2], two, geom) + xs
    return o1 + r2, o2 + r1, ns


def _planes(x, repeat):
    f1, f2 = x[:, :, 0], x[:, :, 1]
    rep = lambda t: t.repeat(1, repeat, 1, 1)
    return f2, rep(f1[:, 0:1]), rep(f1[:, 1:2]), rep(f2[:, 0:1]), rep(f2[:, 1:2])


def _reconstruct(a_o, f2, geom, scale):
    b, h, w = geom
    n_o = from_packed(a_o, b, 2 * scale * scale, h, w)
    return F.pixel_shuffle(n_o, scale) + F.interpolate(f2[:, :2].float(), scale_factor=scale, mode='bilinear',
                                                       align_corners=False)


def _state_in(t, tc):
    return to_packed(_ScaleGrad.apply(t.float(), 1.0 / tc.loss_scale), 128)


def _state_out(a, geom, tc):
    b, h, w = geom
    return _ScaleGrad.apply(from_packed(a, b, 128, h, w), tc.loss_scale)


def forward_plain(model, tc, x, x_h, x_o, init):
    """BMCNet_plain.forward (BMCNet_plain.py:44-68) with autograd."""
    nb = model.neuro
    b, _, _, h, w = x.shape
    geom = (b, h, w)
    sc, rp = model.scale, model.repeat
    k = sc * sc
    x = x.float()
    tc.new_anchor(x.device)
    f2, x1p, x1n, x2p, x2n = _planes(x, rp)
    o = x_o.float() if init else F.pixel_unshuffle(x_o.float(), sc)
    o = _ScaleGrad.apply(o, 1.0 / tc.loss_scale)
    hs = _state_in(x_h, tc)
    in_1, in_2 = torch.cat([x1p, x2p], 1), torch.cat([x1n, x2n], 1)
    n6 = 2 * rp
    # conv_f1(cat[in_1(6), h(128), o1(16)]): sources = h and the packed small planes
    m1 = to_packed(torch.cat([in_1, o[:, :k]], 1), 64)
    m2 = to_packed(torch.cat([in_2, o[:, k:]], 1), 64)
    ms = to_packed(torch.cat([in_1, in_2, o], 1), 64)
    seg_small = _seg(0, n6, 0) + list(range(n6 + 128, n6 + 128 + k))
    seg_small += [-1] * (64 - len(seg_small))
    x1 = conv(tc, nb.conv_f1, [hs, m1], [_seg(n6), seg_small], geom, relu=True)
    x2 = conv(tc, nb.conv_f2, [hs, m2], [_seg(n6), seg_small], geom, relu=True)
    seg_s = list(range(2 * n6)) + list(range(2 * n6 + 128, 2 * n6 + 128 + 2 * k))
    seg_s += [-1] * (64 - len(seg_s))
    xs = conv(tc, nb.conv_fs, [hs, ms], [_seg(2 * n6), seg_s], geom, relu=True)
    if _use_stack(b, h, w):
        rows = b * K.rows_per_image(h, w)
        x12 = torch.cat([x1, x2], 0)
        for blk in nb.para_reschunk:
            x12, xs = bie_stack(tc, blk, x12, xs, 1, geom)
        x1, x2 = x12[:rows], x12[rows:]
    else:
        for blk in nb.para_reschunk:
            x1, x2, xs = bie(tc, blk, x1, x2, xs, geom)
    n_h = conv(tc, nb.conv_h, [xs], [_R128], geom, relu=True)
    a_o = conv(tc, nb.conv_o, [x1, x2], [_R128, _seg(128)], geom)
    pred = _reconstruct(_ScaleGrad.apply(a_o, tc.loss_scale), f2, geom, sc)
    return _state_out(n_h, geom, tc), pred


def forward_full(model, tc, x, x_h, x_h_p, x_h_n, x_o, init):
    """BMCNet.forward (BMCNet.py:95-121) + Backbone.forward (:57-84) + ParallelBlk.forward (:19-32) with autograd."""
    nb = model.neuro
    b, _, _, h, w = x.shape
    geom = (b, h, w)
    sc, rp = model.scale, model.repeat
    k = sc * sc
    x = x.float()
    tc.new_anchor(x.device)
    f2, x1p, x1n, x2p, x2n = _planes(x, rp)
    o = x_o.float() if init else F.pixel_unshuffle(x_o.float(), sc)
    o = _ScaleGrad.apply(o, 1.0 / tc.loss_scale)
    # positional hand-over of the reference: Backbone.forward(xs, hp, hn, hs, o) is called with (x_h, x_h_p, x_h_n)
    hp, hn, hs = _state_in(x_h, tc), _state_in(x_h_p, tc), _state_in(x_h_n, tc)
    n6 = 2 * rp
    mp = to_packed(torch.cat([x1p, x2p, o[:, :k]], 1), 64)           # conv_fpst: cat[xp(6), hp, op(16)]
    mn = to_packed(torch.cat([x1n, x2n, o[:, k:]], 1), 64)
    seg_small = _seg(0, n6, 0) + list(range(n6 + 128, n6 + 128 + k))
    seg_small += [-1] * (64 - len(seg_small))
    xp_st = conv(tc, nb.conv_fpst, [hp, mp], [_seg(n6), seg_small], geom, relu=True)
    xn_st = conv(tc, nb.conv_fnst, [hn, mn], [_seg(n6), seg_small], geom, relu=True)
    m2p, m2n = to_packed(x2p, 64), to_packed(x2n, 64)               # conv_fps: cat[x2p(3), hp]
    seg3 = _seg(0, rp, 64)
    xp_s = conv(tc, nb.conv_fps, [hp, m2p], [_seg(rp), seg3], geom, relu=True)
    xn_s = conv(tc, nb.conv_fns, [hn, m2n], [_seg(rp), seg3], geom, relu=True)
    mo = to_packed(o, 64)                                           # conv_fs: cat[xp_st, xn_st, h*, o(32)]
    seg_o = _seg(384, 2 * k, 64)
    fs = lambda hx: conv(tc, nb.conv_fs, [xp_st, xn_st, hx, mo], [_R128, _seg(128), _seg(256), seg_o], geom, relu=True)
    xs, xs_p, xs_n = fs(hs), fs(hp), fs(hn)
    if _use_stack(b, h, w):
        # ParallelBlk (BMCNet.py:19-32) on stacked operands: S = (xp_s, xp_st, xn_s, xn_st), PN = (xs_p, xs_n).  The four
        # ResidualBlocks are one 4-job launch per convolution, lBIE on the positive and the negative triple shares every
        # launch (same weights, two instances), gBIE is one instance.
        rows = b * K.rows_per_image(h, w)
        S = torch.cat([xp_s, xp_st, xn_s, xn_st], 0)          # (x1, x2) of the positive lBIE instance, then of the negative one
        PN = torch.cat([xs_p, xs_n], 0)
        sl = lambda t, j: t[j * rows:(j + 1) * rows]
        for blk in nb.para_reschunk:
            S = resblock_stack(tc, [blk.conv1, blk.conv1_st, blk.conv2, blk.conv2_st], S, geom)
            X, PN = bie_stack(tc, blk.lBIE, S, PN, 2, geom)                                      # (xp_s, xp_st, xn_s, xn_st)'
            G, xs = bie_stack(tc, blk.gBIE, torch.cat([sl(X, 0), sl(X, 2)], 0), xs, 1, geom)     # (xp_s, xn_s)''
            S = torch.cat([sl(G, 0), sl(X, 1), sl(G, 1), sl(X, 3)], 0)
        hsn = conv_stack(tc, [nb.conv_hs, nb.conv_hp, nb.conv_hn], [torch.cat([xs, PN], 0)], [_R128], [[0, 1, 2]], geom, relu=True)
        n_h, n_hp, n_hn = hsn[:rows], hsn[rows:2 * rows], hsn[2 * rows:]
        a_o = conv(tc, nb.conv_o, [sl(S, 0), sl(S, 2)], [_R128, _seg(128)], geom)
        pred = _reconstruct(_ScaleGrad.apply(a_o, tc.loss_scale), f2, geom, sc)
        return _state_out(n_h, geom, tc), _state_out(n_hp, geom, tc), _state_out(n_hn, geom, tc), pred
    for blk in nb.para_reschunk:
        xp_s = resblock(tc, blk.conv1, xp_s, geom)
        xn_s = resblock(tc, blk.conv2, xn_s, geom)
        xp_st = resblock(tc, blk.conv1_st, xp_st, geom)
        xn_st = resblock(tc, blk.conv2_st, xn_st, geom)
        xp_s, xp_st, xs_p = bie(tc, blk.lBIE, xp_s, xp_st, xs_p, geom)
        xn_s, xn_st, xs_n = bie(tc, blk.lBIE, xn_s, xn_st, xs_n, geom)
        xp_s, xn_s, xs = bie(tc, blk.gBIE, xp_s, xn_s, xs, geom)
    n_h = conv(tc, nb.conv_hs, [xs], [_R128], geom, relu=True)
    n_hp = conv(tc, nb.conv_hp, [xs_p], [_R128], geom, relu=True)
    n_hn = conv(tc, nb.conv_hn, [xs_n], [_R128], geom, relu=True)
    a_o = conv(tc, nb.conv_o, [xp_s, xn_s], [_R128, _seg(128)], geom)
    pred = _reconstruct(_ScaleGrad.apply(a_o, tc.loss_scale), f2, geom, sc)
    return _state_out(n_h, geom, tc), _state_out(n_hp, geom, tc), _state_out(n_hn, geom, tc), pred


_warned_eval_grad = False


def route(model, *tensors):
    """True when this call must record an autograd graph: `model.train()` (train.py:192) with grad mode on.
    In eval() mode the inference kernels run and the outputs carry no grad_fn -- what the reference's own
    inference / validation code asks for (`@torch.no_grad()`, infer_BMCNet.py:144; train.py:476).  A caller who
    left grad mode on in eval() is told once; an INPUT that requires grad cannot be honoured there and raises."""
    global _warned_eval_grad
    if not torch.is_grad_enabled():
        return False
    if model.training:
        return True
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise _lib.BmcError('an input requires grad but the module is in eval() mode: the inference kernels record no '
                            'autograd graph; call .train() for the differentiable path')
    if not _warned_eval_grad and any(p.requires_grad for p in model.parameters()):
        import warnings
        warnings.warn('bmcnet_esr_b200: eval()-mode forward with grad mode on runs the inference kernels and returns '
                      'tensors without grad_fn; use .train() to record an autograd graph (or torch.no_grad() to '
                      'silence this)')
        _warned_eval_grad = True
    return False


def context(model):
    tc = getattr(model, '_train_ctx', None)
    if tc is None or tc.loss_scale != float(model.loss_scale):
        tc = _Ctx(model.loss_scale)
        model._train_ctx = tc
    if len(tc.wcache) > 4096:            # stale versions of the weights (one set per optimiser step)
        tc.wcache.clear()
    return tc


# ------------------------------------------------------------------------------------------ optimiser / collective
def unique_parameters(model):
    """The alias-deduplicated parameters (BMCNet: 2,731,680 elements; plain: 1,003,296)."""
    return list(model.parameters())          # nn.Module.parameters() already yields each shared Parameter once


class FusedAdamAMSGrad:
    """torch.optim.Adam(lr, betas, eps, weight_decay, amsgrad=True) (config/train_nfs.yml:28-34) as ONE kernel over
    flat fp32 buffers (bmc_adam_amsgrad_step).  The parameters are re-homed as views of one flat tensor."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise _lib.BmcError('FusedAdamAMSGrad needs CUDA parameters (no CPU fallback)')
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                p.grad = self.grad[off:off + k].view(p.shape)
                off += k
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.vmax = torch.zeros_like(self.flat)

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:                     # `p.grad = None` by user code would detach the views: restore
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + off * 4:
                p.grad = self.grad[off:off + k].view(p.shape)
            off += k

    def step(self):
        self.step_count += 1
        with torch.cuda.device(self.flat.device):
            check(lib().bmc_adam_amsgrad_step(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                              self.vmax.data_ptr(), self.flat.numel(), self.step_count, self.lr,
                                              self.betas[0], self.betas[1], self.eps, self.weight_decay, stream_ptr()))
        for p in self.params:                     # the kernel wrote through raw pointers: make the change visible
            torch.autograd.graph.increment_version(p)


def allreduce_gradients(opt_or_params, group=None):
    """Data-parallel gradient exchange: ONE all-reduce (NCCL over NVLink) of the alias-deduplicated
    gradients, averaged over ranks.  With FusedAdamAMSGrad the gradients already live in one flat buffer."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if isinstance(opt_or_params, FusedAdamAMSGrad):
        flat = opt_or_params.grad
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        return flat.numel()
    grads = [p.grad for p in opt_or_params if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel()


# ------------------------------------------------------------------------------------------ the iteration as a CUDA graph
class GraphedIteration:
    """One training iteration of the reference's loop (train.py:202-237: zero_grad, the sequence through the model with
    the state carried, `loss += MSELoss`, `loss.backward()`, optional gradient all-reduce, Adam step) with everything
    between `zero_grad` and the end of `backward` recorded ONCE as a CUDA graph and replayed afterwards.

    Why: at the reference's batch (2 per GPU) the eager iteration is host-bound -- ~6,000 kernel-entry calls and ~20,000
    small PyTorch ops per iteration through Python and the autograd engine, 450-480 ms for ~100 ms of GPU work.  Shapes
    are static during training (fixed batch, sequence length and crop), so the launch sequence is too.

    What is inside the graph: zeroing the flat gradient buffer, packing the fp16 weights from the CURRENT fp32
    parameters (forward and mirrored/transposed dgrad packs, once per iteration), all forward / dgrad / wgrad kernels
    and the PyTorch glue.  Outside: the NCCL all-reduce of the flat gradient buffer (world > 1) and the one-launch
    fused Adam step (its bias corrections are host scalars that change every step).

        it = GraphedIteration(model, opt, xs_example, gts_example)
        loss = it(xs, gts)          # copies the inputs into the static buffers, replays, reduces, steps

    `xs`: the sequence of model inputs [B,2,T,H,W] (any strides; `inp_cnt.transpose(1, 2)` in the reference), `gts`: the
    targets [B,2,kH',kW'].  The returned loss is a 0-d tensor that the next call overwrites.

    No Parameter enters the autograd graph of this path (see _ConvFn: the gradient kernels write `.grad` themselves), so
    the capture does not depend on what ran before it -- eager iterations on other streams, loss tensors still alive.
    """

    def __init__(self, model, opt, xs, gts, group=None, warmup=2, capture_error_mode='thread_local', wgrad_stream=True,
                 check_overflow=True):
        if not isinstance(opt, FusedAdamAMSGrad):
            raise _lib.BmcError('GraphedIteration needs FusedAdamAMSGrad (the gradients must live in one static buffer)')
        self.model, self.opt, self.group = model, opt, group
        dev = opt.flat.device
        self.n_state = 3 if hasattr(model.neuro, 'conv_hs') else 1
        self.xs = [x.to(dev).float().contiguous().clone() for x in xs]
        self.gts = [g.to(dev).float().contiguous().clone() for g in gts]
        self.loss = None
        self.overflows = 0            # iterations whose gradients were not finite (step skipped, loss scale halved)
        self.check_overflow = check_overflow
        self._stream = None
        self._capture_args = (capture_error_mode, wgrad_stream)
        model.train()
        self._capture(warmup)

    def _capture(self, warmup):
        model, dev = self.model, self.opt.flat.device
        capture_error_mode, wgrad_stream = self._capture_args
        # Warm-up and capture run on ONE stream of this object.  The autograd engine orders the caller's stream behind the
        # stream every parameter's AccumulateGrad node was created on (even for the undefined gradients this path hands
        # it); a node that survives from the warm-up would otherwise make the capturing stream wait on an uncaptured one
        # ("dependency created on uncaptured work in another stream", seen once under compute-sanitizer).
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        side = self._stream
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                    # lazy module loads, kernel attributes, index caches: not capturable
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        context(model).wcache.clear()                  # the packs must be rebuilt INSIDE the graph, from the live parameters
        tc = context(model)
        tc.side = torch.cuda.Stream(device=dev) if wgrad_stream else None
        self.graph = torch.cuda.CUDAGraph()
        # 'thread_local': a training process has other threads that touch CUDA (pinned-memory loaders, the clock sampler of
        # bench.py); their calls are not part of this stream's capture and must not invalidate it
        try:
            with torch.cuda.graph(self.graph, stream=self._stream, capture_error_mode=capture_error_mode):
                self.loss = self._body()
                if tc.side is not None:
                    torch.cuda.current_stream(dev).wait_stream(tc.side)      # join: the gradients are complete when the graph is
        except RuntimeError as e:
            tc.side = None
            tc.keep.clear()
            if 'uncaptured work' in str(e) or 'during capture' in str(e):
                raise _lib.BmcError(
                    'GraphedIteration: the capture of the training iteration was invalidated (%s): something inside the '
                    'iteration touched a stream that is not part of the capture -- e.g. a custom loss / hook running on '
                    'another stream, or tensors whose autograd history was recorded on another stream.' % str(e).splitlines()[0]) from e
            raise
        tc.side = None
        tc.wcache.clear()                              # (entries point into the graph's pool; never reuse them eagerly)

    def _body(self):
        m, dev = self.model, self.opt.flat.device
        self.opt.zero_grad()
        b, _, _, h, w = self.xs[0].shape
        st = [torch.zeros(b, 128, h, w, device=dev) for _ in range(self.n_state)]
        st.append(torch.zeros(b, 2 * m.scale * m.scale, h, w, device=dev))
        loss, init = 0, True
        for x, gt in zip(self.xs, self.gts):
            st = list(m(x, *st, init))
            init = False
            pred = st[-1]
            if pred.shape[-2:] != gt.shape[-2:]:       # train.py:224-228
                pred = F.interpolate(pred, size=gt.shape[-2:], mode='bicubic', align_corners=False)
            loss = loss + F.mse_loss(pred, gt)
        loss.backward()
        context(m).keep.clear()
        return loss.detach()

    def __call__(self, xs=None, gts=None):
        if xs is not None:
            for dst, src in zip(self.xs, xs):
                dst.copy_(src, non_blocking=True)
        if gts is not None:
            for dst, src in zip(self.gts, gts):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                      and torch.distributed.get_world_size() > 1):
            allreduce_gradients(self.opt, self.group)
        # The 16-bit activation gradients run under a static loss scale; a gradient spike (BPTT over the sequence does
        # produce them: fp32 autograd on the same data shows max |grad| jump from 0.6 to 94 for one iteration) overflows
        # them.  Like torch.amp's GradScaler: a non-finite gradient buffer skips the step and halves the scale -- which
        # is a constant of the captured graph, so the iteration is re-captured.  After the all-reduce every rank sees the
        # same buffer and takes the same decision.
        if self.check_overflow and not bool(torch.isfinite(self.opt.grad).all()):
            self.overflows += 1
            if self.model.loss_scale > 1.0:
                self.model.loss_scale = float(self.model.loss_scale) / 2.0
                loss = self.loss.clone()
                self.graph = None
                self._capture(1)
                return loss
            return self.loss
        self.opt.step()
        return self.loss
