"""Thin host wrappers over the per-kernel C-ABI entry points (bmc_conv_gemm, bmc_attention_weights,
bmc_pack_nchw, ...).  Used by the sub-module forwards in models/submodules.py and by the unit
parity tests; the full models go through bmc_model_forward instead.

Layout reminder: activations are act16 [B*R, C] with R = roundup((H+2)*(W+2), 128);
row r of image b is padded pixel (r // (W+2), r % (W+2)); halo and tail rows are zero.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmJob, check, lib, stream_ptr


def rows_per_image(h, w):
    return ((h + 2) * (w + 2) + 127) // 128 * 128


def _need_cuda(*ts):
    """All operands on ONE CUDA device; returns a context that makes it the current device, so that
    `stream_ptr()` is that device's current stream and the launch lands there."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.BmcError('bmcnet_esr_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor'
                                % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _lib.BmcError('operands on different devices: %s and %s' % (dev, t.device))
    return torch.cuda.device(dev)


def pack_nchw(x, c_pad=None):
    """fp32 [B,C,H,W] -> padded NHWC act16 [B*R, c_pad] (zero halo)."""
    b, c, h, w = x.shape
    c_pad = c_pad or (c + 63) // 64 * 64
    with _need_cuda(x):
        out = torch.zeros(b * rows_per_image(h, w), c_pad, dtype=_lib.act_dtype(), device=x.device)
        x = x.contiguous().float()
        check(lib().bmc_pack_nchw(x.data_ptr(), b, c, h, w, out.data_ptr(), c_pad, 0, stream_ptr()))
    return out


def unpack_nchw(a, b, c, h, w):
    """padded NHWC act16 [B*R, c_pad] -> fp32 [B,C,H,W]."""
    with _need_cuda(a):
        out = torch.empty(b, c, h, w, dtype=torch.float32, device=a.device)
        check(lib().bmc_unpack_nchw(a.data_ptr(), b, c, h, w, a.shape[1], 0, out.data_ptr(), stream_ptr()))
    return out


def pack_conv_weight(weight, seg_channels):
    """Conv2d weight [N, Cin, k, k] -> act16 chunk-major [K/64, N, 64], K order (segment, tap, channel).

    `seg_channels`: list of (first input channel, count) per concatenated source; each segment is
    zero-padded to a multiple of 64 channels."""
    n, cin, kh, kw = weight.shape
    taps = kh * kw
    cols = []
    for first, cnt in seg_channels:
        pad = (cnt + 63) // 64 * 64
        seg = weight[:, first:first + cnt].reshape(n, cnt, taps).permute(0, 2, 1)      # [N, taps, cnt]
        seg = torch.nn.functional.pad(seg, (0, pad - cnt))
        cols.append(seg.reshape(n, taps * pad))
    wk = torch.cat(cols, 1)                                                            # [N, K]
    k = wk.shape[1]
    return wk.reshape(n, k // 64, 64).permute(1, 0, 2).contiguous().to(_lib.act_dtype())


def conv_gemm(srcs, wpk, bias, b, h, w, taps, n=128, relu=False, residual=None, ln=None, impl=0,
              out_f32=False):
    """One conv-gemm job (see include/bmc_b200.h: bmc_conv_gemm).  srcs: packed act16 sources."""
    rows = b * rows_per_image(h, w)
    dev = srcs[0].device
    out = torch.empty(rows, n, dtype=_lib.act_dtype(), device=dev)
    outf = torch.empty(rows, n, dtype=torch.float32, device=dev) if out_f32 else None
    j = GemmJob()
    j.n_seg = len(srcs)
    for i, s in enumerate(srcs):
        j.a[i] = s.data_ptr(); j.a_rows[i] = s.shape[0]; j.a_ch[i] = s.shape[1]; j.a_row_base[i] = 0
    j.w = wpk.data_ptr(); j.w_rows = wpk.shape[1]; j.w_k = wpk.shape[0] * 64
    j.w_row_base = 0; j.w_img_stride = 0
    bias = None if bias is None else bias.contiguous().float()
    j.bias = bias.data_ptr() if bias is not None else None
    j.residual = residual.data_ptr() if residual is not None else None
    j.out_act16 = out.data_ptr(); j.out_f32 = outf.data_ptr() if out_f32 else None
    j.relu = int(relu)
    keep = []
    if ln is not None:
        g, bt, eps = ln
        g, bt = g.contiguous().float(), bt.contiguous().float()
        keep += [g, bt]
        j.ln_gamma = g.data_ptr(); j.ln_beta = bt.data_ptr(); j.ln_eps = eps
    with _need_cuda(*srcs, wpk, bias, residual):
        check(lib().bmc_conv_gemm(C.byref(j), 1, n, taps, b, h, w, impl, stream_ptr()))
    return (out, outf) if out_f32 else out


def attention_weights(centres, v, b, h, w, scale, n_split=4, impl=0):
    """softmax(centres^T v * scale) per image -> act16 [B, 2, 128, 64] chunk-major dynamic weights."""
    dev = centres.device
    partial = torch.empty(b, n_split, 128, 128, dtype=torch.float32, device=dev)
    probs = torch.empty(b * 256, 64, dtype=_lib.act_dtype(), device=dev)
    with _need_cuda(centres, v):
        check(lib().bmc_attention_weights(centres.data_ptr(), v.data_ptr(), b, h, w, scale, partial.data_ptr(),
                                          n_split, probs.data_ptr(), impl, stream_ptr()))
    return probs, partial


def apply_dynamic_weights(v, probs, b, h, w, residual=None, impl=0):
    """out[b] = v[b] . P[b]^T (+ residual): the `softmax(att) @ v` product of BIE."""
    rows = b * rows_per_image(h, w)
    out = torch.empty(rows, 128, dtype=_lib.act_dtype(), device=v.device)
    j = GemmJob()
    j.n_seg = 1
    j.a[0] = v.data_ptr(); j.a_rows[0] = v.shape[0]; j.a_ch[0] = 128; j.a_row_base[0] = 0
    j.w = probs.data_ptr(); j.w_rows = 128; j.w_k = 128; j.w_row_base = 0; j.w_img_stride = 256
    j.residual = residual.data_ptr() if residual is not None else None
    j.out_act16 = out.data_ptr()
    with _need_cuda(v, probs, residual):
        check(lib().bmc_conv_gemm(C.byref(j), 1, 128, 1, b, h, w, impl, stream_ptr()))
    return out


def layernorm_rows(a, gamma, beta, eps):
    out = torch.empty_like(a)
    g, bt = gamma.contiguous().float(), beta.contiguous().float()
    with _need_cuda(a, g, bt):
        check(lib().bmc_layernorm_rows(a.data_ptr(), g.data_ptr(), bt.data_ptr(), eps, a.shape[0], out.data_ptr(),
                                       stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------ training kernels
def relu_backward(dy, y):
    """dx = dy * (y > 0) on act16 tensors (bmc_relu_backward)."""
    dx = torch.empty_like(dy)
    with _need_cuda(dy, y):
        check(lib().bmc_relu_backward(dy.data_ptr(), y.data_ptr(), dy.numel(), dx.data_ptr(), stream_ptr()))
    return dx


def conv_wgrad(dy, x, taps, b, h, w, cmap, cin_total, n_out, scale, grad_w, grad_b, workspace, n_split):
    """grad_w[co][cmap[ci]][tap] += scale * sum_rows dy[row][co] * x[row + off_tap][ci] (and grad_b += scale * column
    sums of dy when given): bmc_conv_wgrad.  dy: act16 [B*R,128]; x: act16 [B*R, 64|128]; cmap: int32 [x_ch]."""
    with _need_cuda(dy, x, cmap, grad_w, grad_b, workspace):
        check(lib().bmc_conv_wgrad(dy.data_ptr(), x.data_ptr(), x.shape[1], taps, b, h, w, cmap.data_ptr(), cin_total,
                                   n_out, scale, grad_w.data_ptr(), grad_b.data_ptr() if grad_b is not None else None,
                                   workspace.data_ptr(), workspace.numel(), n_split, stream_ptr()))


def layernorm_rows_backward(x, dy, gamma, eps, scale, grad_gamma, grad_beta, workspace):
    """dx of the channel LayerNorm (bmc_layernorm_rows_backward); grad_gamma / grad_beta (fp32 [128]) += scale * sums."""
    dx = torch.empty_like(dy)
    g = gamma.contiguous().float()
    with _need_cuda(x, dy, g, grad_gamma, grad_beta, workspace):
        check(lib().bmc_layernorm_rows_backward(x.data_ptr(), dy.data_ptr(), g.data_ptr(), eps, x.shape[0], dx.data_ptr(), scale,
                                                grad_gamma.data_ptr(), grad_beta.data_ptr(), workspace.data_ptr(),
                                                workspace.numel(), stream_ptr()))
    return dx


def conv_gemm_stack(srcs, slabs, wstack, biases, n_jobs, b, h, w, taps, relu=False):
    """J = n_jobs convolutions of one shape in ONE launch (bmc_conv_gemm with n_jobs jobs) on STACKED operands, so that
    the launch needs three or four tensor maps whatever J is:
      srcs[s]    packed act16 [n_slabs_s * rows, C_s]: segment s of job j reads slab slabs[s][j] (rows = B * R)
      wstack     act16 [K/64, J*128, 64]: the J chunk-major weight matrices side by side (job j = rows j*128 .. j*128+127)
      biases     list of J fp32 [128] tensors, or None
    Returns act16 [J * rows, 128], job j in slab j."""
    rows = b * rows_per_image(h, w)
    dev = srcs[0].device
    out = torch.empty(n_jobs * rows, 128, dtype=_lib.act_dtype(), device=dev)
    jobs = (GemmJob * n_jobs)()
    keep = []
    for jn in range(n_jobs):
        j = jobs[jn]
        j.n_seg = len(srcs)
        for i, s in enumerate(srcs):
            j.a[i] = s.data_ptr(); j.a_rows[i] = s.shape[0]; j.a_ch[i] = s.shape[1]; j.a_row_base[i] = slabs[i][jn] * rows
        j.w = wstack.data_ptr(); j.w_rows = wstack.shape[1]; j.w_k = wstack.shape[0] * 64
        j.w_row_base = jn * 128; j.w_img_stride = 0
        if biases is not None:
            bt = biases[jn].contiguous().float()
            keep.append(bt)
            j.bias = bt.data_ptr()
        j.out_act16 = out.data_ptr(); j.out_row_base = jn * rows
        j.relu = int(relu)
    with _need_cuda(*srcs, wstack, *keep):
        check(lib().bmc_conv_gemm(jobs, n_jobs, 128, taps, b, h, w, 0, stream_ptr()))
    return out
