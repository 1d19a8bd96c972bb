"""The reference's OWN `load_model` + `infer_body` (infer_BMCNet.py:20-116, infer_BMCNet_plain.py) executed from the
unmodified tree staged under baseline/_ref, twice:

  * as shipped (reference modules, fp32, on the CPU) -- the oracle run;
  * with ONLY the model import line swapped to bmcnet_esr_b200 on cuda:0.

Everything between `torch.load` and the per-frame MSE is the reference's code in both runs: `load_model`
(strict `load_state_dict`, `.to(device)`, `.eval()`), the zero-state construction, the recurrent hand-back of
`h, hp, hn, prediction`, the `.transpose(1, 2)` input view, the bicubic resize and `nn.MSELoss`.

Environment shims (none on the hot path): the HDF5 dataloader class is replaced by a synthetic in-memory loader
with the same item structure (h5py and datasets are absent), `MetricTracker` by a dict-based stand-in (the
reference's pandas bookkeeping, myutils/utils.py:84-106, raises under pandas 3 copy-on-write), and the
matplotlib / skimage / cv2 imports by inert stubs (oracle/reference_tree.py).

Bars: every frame's prediction max-abs <= 1e-2 vs the reference run, esr_mse within 1e-3 relative, bicubic_mse equal
to 1e-6 relative (same CPU code in both runs)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import bmcnet_fp32 as O
from oracle import reference_tree as R
from oracle.make_golden import synth_counts

pytestmark = pytest.mark.gpu


class _Tracker:
    def __init__(self, keys, writer=None):
        self.tot = {k: 0.0 for k in keys}
        self.cnt = {k: 0 for k in keys}

    def reset(self):
        for k in self.tot:
            self.tot[k], self.cnt[k] = 0.0, 0

    def update(self, key, value, n=1):
        self.tot[key] += value * n
        self.cnt[key] += n

    def result(self):
        return {k: (self.tot[k] / self.cnt[k] if self.cnt[k] else 0.0) for k in self.tot}


class _Logger:
    def log_dict(self, d, name):
        pass

    def log_info(self, s):
        pass


def _fake_loader_cls(frames, h, w, gh, gw, seed):
    """Items shaped like InferenceHDF5DataLoaderSequence's (h5dataloader.py:293-330): a list of dicts with
    inp_cnt [1, seqn=2, 2, H, W] and gt_cnt [1, 2, 2, gH, gW]."""
    g = torch.Generator().manual_seed(seed)
    items = []
    for i in range(frames):
        x = synth_counts(1, h, w, seed + i)                         # [1, 2(pol), 2(T), H, W]
        inp = x.transpose(1, 2).contiguous()                        # [1, T, pol, H, W]
        gt = torch.poisson(torch.full((1, 2, 2, gh, gw), 0.02), generator=g)
        items.append([{'inp_cnt': inp, 'gt_cnt': gt}])

    class Loader:
        def __init__(self, data_path, config):
            self.dataset = types.SimpleNamespace(gt_sensor_resolution=(gh, gw), inp_sensor_resolution=(h, w))

        def __len__(self):
            return len(items)

        def __iter__(self):
            return iter(items)

    return Loader


def _run(script, swap, ckpt_path, device, loader_cls, tmp, tag):
    mod = R.load_script(script, swap=swap)
    mod.InferenceHDF5DataLoaderSequence = loader_cls
    mod.MetricTracker = _Tracker
    mod.tqdm = lambda it, total=None: it
    cfg = {'dataset': {'scale': 4, 'sequence': {'sequence_length': 9, 'seqn': 2}}}
    model = mod.load_model(cfg, ckpt_path, device)
    rec = {'pred': [], 'esr': [], 'bic': []}
    mse = torch.nn.MSELoss()

    def rec_mse(a, b):
        v = mse(a, b)
        if len(rec['esr']) == len(rec['bic']):
            rec['pred'].append(a.clone())
            rec['esr'].append(v.item())
        else:
            rec['bic'].append(v.item())
        return v

    vis = types.SimpleNamespace(plot_event_cnt=lambda *a, **k: None)
    with torch.no_grad():
        res = mod.infer_body(cfg, 'synthetic', model, os.path.join(tmp, 'img_' + tag), _Logger(), device, vis,
                             {'mse': rec_mse})
    return model, rec, res


@pytest.mark.skipif(not R.available(), reason='baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)')
@pytest.mark.parametrize('kind', ['plain', 'full'])
def test_reference_infer_body_with_only_the_import_swapped(kind, tmp_path, plain_ckpt):
    frames = 8
    if kind == 'plain':
        script = 'infer_BMCNet_plain.py'
        swap = ('from models.BMCNet_plain import BMCNet_plain', 'from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain')
        ckpt = os.path.join(R.REF_ROOT, 'pretrain', 'BMCNet_plain_nfs_x4.pth')          # the shipped file itself
        h, w, gh, gw = 22, 40, 90, 160                                                    # scripts/infer_ours.sh:12 (down16): resize path
    else:
        script = 'infer_BMCNet.py'
        swap = ('from models.BMCNet import BMCNet', 'from bmcnet_esr_b200.models.BMCNet import BMCNet')
        ckpt = str(tmp_path / 'bmcnet_surrogate.pth')                                     # BMCNet checkpoints are not shipped
        torch.save(O.surrogate_state_dict(plain=False, seed=11, transplant=plain_ckpt), ckpt)
        h, w, gh, gw = 31, 56, 124, 222                                                   # EventZoom: 124x224 -> 124x222
    loader = _fake_loader_cls(frames, h, w, gh, gw, seed=4000)
    ref_model, ref, ref_res = _run(script, None, ckpt, torch.device('cpu'), loader, str(tmp_path), 'ref')
    got_model, got, got_res = _run(script, swap, ckpt, torch.device('cuda:0'), loader, str(tmp_path), 'b200')
    assert type(ref_model).__module__.startswith('models.') and type(got_model).__module__.startswith('bmcnet_esr_b200.')
    assert list(ref_model.state_dict().keys()) == list(got_model.state_dict().keys())
    assert len(ref['pred']) == len(got['pred']) == frames
    for i in range(frames):
        err = (ref['pred'][i] - got['pred'][i]).abs().max().item()
        assert err <= 1e-2, (kind, i, err)
        assert abs(got['esr'][i] - ref['esr'][i]) <= 1e-3 * ref['esr'][i], (kind, i, got['esr'][i], ref['esr'][i])
        assert abs(got['bic'][i] - ref['bic'][i]) <= 1e-6 * ref['bic'][i]
    assert abs(got_res['esr_mse'] - ref_res['esr_mse']) <= 1e-3 * ref_res['esr_mse']
    assert np.isclose(got_res['params'], ref_res['params'])                               # infer_BMCNet.py:70-72
