"""GPU parity of the evaluation tail (csrc/eval_tail.cu via bmcnet_esr_b200.metrics.sr_metrics) against
the CPU oracle, which makes the reference's own torch calls (infer_BMCNet.py:77-87).
Bar: 1e-5 relative on both means (fp32 bicubic taps summed in another order; the sums are kept in fp64)."""
import pytest
import torch

from oracle import eval_tail as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _case(b, h, w, hp, wp, hg, wg, seed):
    g = torch.Generator().manual_seed(seed)
    inp = torch.randint(0, 4, (b, 2, h, w), generator=g).float()            # LR event counts
    pred = torch.rand(b, 2, hp, wp, generator=g) * 3 - 0.5                  # SR prediction
    gt = torch.randint(0, 3, (b, 2, hg, wg), generator=g).float()           # HR event counts
    return pred, inp, gt


@pytest.mark.parametrize('shape', [
    (2, 45, 80, 180, 320, 180, 320),     # NFS x4: prediction already at the ground-truth size
    (3, 31, 56, 124, 224, 124, 222),     # EventZoom: 124x224 prediction resized to the 124x222 ground truth
    (1, 22, 40, 88, 160, 90, 160),       # NFS down16: 88x160 -> 90x160 (scripts/infer_ours.sh)
    (1, 5, 7, 20, 28, 19, 30),           # tiny, both directions at once
])
def test_sr_metrics_match_reference_calls(shape):
    from bmcnet_esr_b200.metrics import sr_metrics
    b, h, w, hp, wp, hg, wg = shape
    pred, inp, gt = _case(b, h, w, hp, wp, hg, wg, seed=sum(shape))
    ref = O.sr_metrics(pred, inp, gt)
    got = sr_metrics(pred.cuda(), inp.cuda(), gt.cuda())
    assert got[0].is_cuda and got[0].dim() == 0
    for r, g_ in zip(ref, got):
        assert abs(g_.item() - r) <= RTOL * max(abs(r), 1e-6), (g_.item(), r)


def test_sr_metrics_identity_is_zero():
    from bmcnet_esr_b200.metrics import sr_metrics
    pred, inp, gt = _case(2, 8, 8, 32, 32, 32, 32, seed=5)
    e, _ = sr_metrics(gt.cuda(), inp.cuda(), gt.cuda())
    assert e.item() == 0.0


def test_sr_metrics_rejects_cpu_tensors():
    from bmcnet_esr_b200 import _lib
    from bmcnet_esr_b200.metrics import sr_metrics
    pred, inp, gt = _case(1, 4, 4, 16, 16, 16, 16, seed=1)
    with pytest.raises(_lib.BmcError):
        sr_metrics(pred, inp, gt)
