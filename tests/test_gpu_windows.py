"""GPU parity of the dataloader window pipeline on raw recordings (bmcnet_esr_b200/dataloader/h5windows.py,
csrc/encode.cu) against the CPU oracle (oracle/h5windows_np.py) and the reference-generated golden:
bit-exact (integer counts, float32 casts and one float32 division)."""
import os

import numpy as np
import pytest
import torch

from oracle import h5windows_np as O
from oracle.make_golden import synth_recording

pytestmark = pytest.mark.gpu


def _cuda(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def test_event_formatting_matches_reference_golden(golden_dir):
    from bmcnet_esr_b200.dataloader.h5windows import event_formatting
    d = np.load(os.path.join(golden_dir, 'fmt_events.npz'))
    for name in ('w2048', 'rel2048', 'w5', 'w1'):
        got = event_formatting(_cuda(d[name + '_xs'], d[name + '_ys'], d[name + '_ts'], d[name + '_ps']))
        assert got.dtype == torch.float32 and tuple(got.shape) == d[name + '_out'].shape
        assert np.array_equal(got.cpu().numpy(), d[name + '_out'], equal_nan=True), name


@pytest.mark.parametrize('n,h,w,window,sliding,oor', [
    (50000, 45, 80, 2048, 1024, 0.0),      # NFS: window 2048, sliding 1024 (config/train_nfs.yml:7-10)
    (20000, 31, 56, 1024, 512, 0.03),      # EventZoom, with out-of-range events (F9 leak into row H-1, col 0)
    (2100, 22, 40, 2048, 1024, 0.0),       # last window clipped to n - 1
    (300, 9, 11, 2048, 1024, 0.1),         # recording shorter than one stride: no windows
])
def test_windows_to_counts_bit_exact(n, h, w, window, sliding, oor):
    from bmcnet_esr_b200.dataloader.h5windows import windows_to_counts
    xs, ys, ts, ps = synth_recording(n, h, w, seed=n, oor=oor, t_base=0.0)
    ref = O.windows_to_counts(xs, ys, ts, ps, window, sliding, (h, w))
    cx, cy, cp = _cuda(xs, ys, ps)
    got = windows_to_counts(cx, cy, cp, window, sliding, (h, w))
    assert tuple(got.shape) == ref.shape
    assert np.array_equal(got.cpu().numpy(), ref)
    if len(ref):
        # every event of a window is counted once (both polarities), wherever it lands
        k = O.compute_k_indices(n, window, sliding)
        assert got[0].sum().item() <= k[0][1] - k[0][0]


def test_windows_feed_the_model_like_the_dataloader():
    """[w_i, w_{i+1}] pairs of consecutive windows are the model's input (infer_BMCNet.py:48-50)."""
    from bmcnet_esr_b200.dataloader.h5windows import windows_to_counts
    xs, ys, ts, ps = synth_recording(12000, 22, 40, seed=3, t_base=0.0)
    cnt = windows_to_counts(*_cuda(xs, ys, ps), 2048, 1024, (22, 40))          # [n_win, 2, H, W]
    x = torch.stack([cnt[:-1], cnt[1:]], 1).transpose(1, 2)                    # [n_win-1, 2, T=2, H, W]
    assert tuple(x.shape) == (cnt.shape[0] - 1, 2, 2, 22, 40)
    assert torch.equal(x[0, :, 1], cnt[1])


def test_raw_recordings_must_be_cuda_int16_float64():
    from bmcnet_esr_b200 import _lib
    from bmcnet_esr_b200.dataloader.h5windows import windows_to_counts
    with pytest.raises(_lib.BmcError):
        windows_to_counts(torch.zeros(8, dtype=torch.int16), torch.zeros(8, dtype=torch.int16),
                          torch.zeros(8, dtype=torch.float64), 4, 2, (4, 4))
