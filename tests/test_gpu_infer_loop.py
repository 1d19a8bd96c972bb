"""End to end: the reference's inference loop (infer_BMCNet_plain.py / infer_BMCNet.py:46-87) on a synthetic raw
recording, once with the device pipeline (windows_to_counts -> BMCNet_plain -> sr_metrics, nothing leaves the GPU
inside the loop) and once with the CPU oracle restatement of every stage (window indexing, event_formatting,
events_to_channels, fp32 forward, bicubic + MSE).  Shipped checkpoint, NFS down16 shape (LR 22x40, GT 90x160 so that
the prediction 88x160 is bicubic-resized like scripts/infer_ours.sh:12).

Bars: count frames bit-exact; per-frame esr_mse / bicubic_mse within 1e-3 relative (the SR output itself is held to
max-abs 1e-2 by test_gpu_model.py; the MSE against ~Poisson ground truth is dominated by the ground truth)."""
import numpy as np
import pytest
import torch

from oracle import bmcnet_fp32 as OM
from oracle import eval_tail as OE
from oracle import h5windows_np as OW
from oracle.make_golden import synth_recording

pytestmark = pytest.mark.gpu


def test_inference_loop_matches_cpu_restatement(plain_ckpt):
    from bmcnet_esr_b200.dataloader.h5windows import sequence_tuples, windows_to_counts
    from bmcnet_esr_b200.metrics import sr_metrics
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    h, w, gh, gw = 22, 40, 90, 160
    window, sliding = 2048, 1024
    n_frames = 5
    n_lr = (n_frames + 2) * (window - sliding) + window
    xs, ys, ts, ps = synth_recording(n_lr, h, w, seed=5, oor=0.01, t_base=0.0)
    gxs, gys, gts, gps = synth_recording(16 * n_lr, gh, gw, seed=6, t_base=0.0)      # the HR stream has 16x the events
    # ---- CPU restatement
    inp_ref = OW.windows_to_counts(xs, ys, ts, ps, window, sliding, (h, w))
    gt_ref = OW.windows_to_counts(gxs, gys, gts, gps, 16 * window, 16 * sliding, (gh, gw))
    st = [torch.zeros(1, 128, h, w), torch.zeros(1, 32, h, w)]
    ref = []
    for i in range(n_frames):
        x = torch.from_numpy(np.stack([inp_ref[i], inp_ref[i + 1]]))[None].transpose(1, 2)       # infer_BMCNet.py:48-50
        st = list(OM.bmcnet_plain_forward(plain_ckpt, x, *st, i == 0))
        ref.append(OE.sr_metrics(st[-1], torch.from_numpy(inp_ref[i + 1])[None], torch.from_numpy(gt_ref[i + 1])[None]))
    # ---- device pipeline
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    inp = windows_to_counts(cu(xs), cu(ys), cu(ps), window, sliding, (h, w))
    gt = windows_to_counts(cu(gxs), cu(gys), cu(gps), 16 * window, 16 * sliding, (gh, gw))
    assert np.array_equal(inp.cpu().numpy(), inp_ref) and np.array_equal(gt.cpu().numpy(), gt_ref)
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    m = m.cuda().eval()
    tuples = sequence_tuples(inp, 2)                                                 # [n_win - 1, 2, 2, H, W] view
    hs, o = torch.zeros(1, 128, h, w).cuda(), torch.zeros(1, 32, h, w).cuda()
    got = []
    for i in range(n_frames):
        hs, o = m(tuples[i:i + 1].transpose(1, 2), hs, o, i == 0)
        got.append(sr_metrics(o, inp[i + 1][None], gt[i + 1][None]))                 # 0-dim CUDA tensors, no sync here
    for i, ((ge, gb), (re, rb)) in enumerate(zip(got, ref)):
        assert abs(gb.item() - rb) <= 1e-5 * rb, (i, 'bicubic_mse', gb.item(), rb)
        assert abs(ge.item() - re) <= 1e-3 * re, (i, 'esr_mse', ge.item(), re)
