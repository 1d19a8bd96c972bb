"""Evaluation tail of the reference's inference loop (infer_BMCNet.py:77-87) on the device.

The reference copies every prediction to the host, bicubic-resizes it to the ground-truth size when the
two differ, bicubic-upsamples the LR count frame as a baseline and takes nn.MSELoss of both against the
ground truth.  `sr_metrics` does all of that in one CUDA kernel (csrc/eval_tail.cu) and returns the two
means as 0-dim CUDA tensors, so the loop needs no per-frame device->host synchronisation."""
import ctypes as C

import torch

from . import _lib

__all__ = ['sr_metrics']


def sr_metrics(prediction, inp_cnt, gt_cnt):
    """(esr_mse, bicubic_mse) of `prediction` [B,2,Hp,Wp] and of the bicubic upsampling of `inp_cnt`
    [B,2,H,W] against `gt_cnt` [B,2,Hg,Wg] -- infer_BMCNet.py:77-84.  CUDA float32 tensors only."""
    for t in (prediction, inp_cnt, gt_cnt):
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4:
            raise _lib.BmcError('sr_metrics needs 4-D CUDA float32 tensors (no CPU fallback)')
    if not (prediction.shape[:2] == inp_cnt.shape[:2] == gt_cnt.shape[:2]):
        raise _lib.BmcError('sr_metrics: batch / channel sizes differ: %s %s %s' % (
            tuple(prediction.shape), tuple(inp_cnt.shape), tuple(gt_cnt.shape)))
    pred, inp, gt = prediction.contiguous(), inp_cnt.contiguous(), gt_cnt.contiguous()
    b, c, hp, wp = pred.shape
    h, w = inp.shape[2:]
    hg, wg = gt.shape[2:]
    sums = torch.empty(2, dtype=torch.float64, device=gt.device)
    with torch.cuda.device(gt.device):
        _lib.check(_lib.lib().bmc_sr_metrics(C.c_void_p(pred.data_ptr()), b, c, hp, wp, C.c_void_p(inp.data_ptr()), h, w,
                                             C.c_void_p(gt.data_ptr()), hg, wg, C.c_void_p(sums.data_ptr()),
                                             _lib.stream_ptr()))
    mse = (sums / float(gt.numel())).to(torch.float32)
    return mse[0], mse[1]
