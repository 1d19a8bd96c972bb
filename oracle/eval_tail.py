"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product) of the evaluation tail of the
reference's inference loop, infer_BMCNet.py:77-87: the same torch CPU calls the reference makes
(third-party arithmetic: ATen upsample_bicubic2d + mse_loss, torch version recorded by the caller)."""
import torch
import torch.nn.functional as f


def sr_metrics(prediction, inp_cnt, gt_cnt):
    """(esr_mse, bicubic_mse) as python floats -- infer_BMCNet.py:77-87 line by line."""
    mse = torch.nn.MSELoss()                                          # infer_BMCNet.py:244-246
    esr_cnt = prediction.cpu()                                        # :77
    gt_cnt = gt_cnt.cpu()
    if esr_cnt.size()[-2:] != gt_cnt.size()[-2:]:                     # :78-79
        esr_cnt = f.interpolate(esr_cnt, size=gt_cnt.size()[-2:], mode='bicubic', align_corners=False)
    bicubic_cnt = f.interpolate(inp_cnt.cpu(), size=tuple(gt_cnt.size()[-2:]), mode='bicubic', align_corners=False)   # :80
    return mse(esr_cnt, gt_cnt).item(), mse(bicubic_cnt, gt_cnt).item()                                                # :83-84
