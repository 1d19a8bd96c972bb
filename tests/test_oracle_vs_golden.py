"""Pin the CPU oracle (oracle/) against golden vectors produced by the reference itself.

The fixtures in tests/golden/ were written by oracle/make_golden.py, which imports the
reference (`dataloader/encodings.py`, `models/*.py`) in the build container.  CPU only.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import bmcnet_fp32 as M
from oracle import encodings_np as E


def _enc_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, 'enc_*.npz')))


ENC_FNS = {
    'image': lambda a, h, w, B: E.events_to_image(a[0], a[1], a[3], sensor_size=(h, w)),
    'channels': lambda a, h, w, B: E.events_to_channels(a[0], a[1], a[3], sensor_size=(h, w)),
    'voxel': lambda a, h, w, B: E.events_to_voxel(a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'image_torch': lambda a, h, w, B: E.events_to_image_torch(a[0], a[1], a[3], sensor_size=(h, w)),
    'image_torch_bilinear': lambda a, h, w, B: E.events_to_image_torch(
        a[0], a[1], a[3], sensor_size=(h, w), interpolation='bilinear'),
    'stack_polarity': lambda a, h, w, B: E.events_to_stack_polarity(
        a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'stack_no_polarity': lambda a, h, w, B: E.events_to_stack_no_polarity(
        a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'voxel_torch': lambda a, h, w, B: E.events_to_voxel_torch(
        a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'voxel_torch_hard': lambda a, h, w, B: E.events_to_voxel_torch(
        a[0], a[1], a[2], a[3], B, sensor_size=(h, w), temporal_bilinear=False),
}


@pytest.mark.parametrize('fname', sorted(ENC_FNS))
def test_encoder_oracle_bit_exact_vs_reference_goldens(golden_dir, fname):
    cases = _enc_cases(golden_dir)
    assert len(cases) >= 8
    for path in cases:
        g = np.load(path)
        h, w, B = int(g['h']), int(g['w']), int(g['B'])
        args = [g['in_' + k].copy() for k in ('xs', 'ys', 'ts', 'ps')]
        out = ENC_FNS[fname](args, h, w, B)
        ref = g['out_' + fname]
        assert out.shape == ref.shape, (path, fname)
        # the oracle keeps the reference's serial fp32 order, so even weighted sums are bit-exact
        assert np.array_equal(out, ref), (path, fname, float(np.abs(out - ref).max()))
        for i, k in enumerate(('xs', 'ys', 'ts', 'ps')):          # in-place side effects (F9)
            key = 'mut_%s_%s' % (fname, k)
            want = g[key] if key in g.files else g['in_' + k]
            assert np.array_equal(args[i], want), (path, fname, k)


def test_binary_search_returns_any_equal_index():
    # the search probes l, r, mid each round and returns the first hit
    t = np.array([0.0, 0.1, 0.1, 0.1, 0.1, 0.5, 0.9], dtype=np.float32)
    assert E.binary_search_torch_tensor(t, 0, len(t) - 1, np.float32(0.1)) == 3
    assert E.binary_search_torch_tensor(t, 0, len(t) - 1, np.float32(0.3)) == 5
    assert E.binary_search_torch_tensor(t, 0, len(t) - 1, np.float32(0.3), side='right') == 4
    assert E.binary_search_torch_tensor(t, 0, len(t) - 1, np.float32(2.0), side='right') == 6
    assert E.binary_search_torch_tensor(t, 0, len(t) - 1, np.float32(-1.0), side='right') == -1


def _rollout(fwd, sd, g, n_state):
    x = torch.from_numpy(g['x'])
    b, h, w = x.shape[1], x.shape[-2], x.shape[-1]
    st = [torch.zeros(b, 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w)]
    outs = []
    init = True
    for s in range(x.shape[0]):
        st = list(fwd(sd, x[s], *st, init))
        init = False
        outs.append(st[-1])
    return torch.stack(outs), st[:-1]


def test_plain_oracle_vs_reference_shipped_checkpoint(golden_dir, plain_ckpt):
    g = np.load(os.path.join(golden_dir, 'model_plain_shipped.npz'))
    outs, hid = _rollout(M.bmcnet_plain_forward, plain_ckpt, g, 2)
    assert torch.allclose(outs, torch.from_numpy(g['x_o']), atol=1e-5, rtol=1e-5)
    assert torch.allclose(hid[0], torch.from_numpy(g['x_h']), atol=1e-5, rtol=1e-5)


def test_plain_oracle_vs_reference_surrogate(golden_dir):
    g = np.load(os.path.join(golden_dir, 'model_plain_surrogate.npz'))
    sd = M.surrogate_state_dict(plain=True, seed=int(g['seed']))
    outs, hid = _rollout(M.bmcnet_plain_forward, sd, g, 2)
    assert torch.allclose(outs, torch.from_numpy(g['x_o']), atol=1e-5, rtol=1e-5)
    assert torch.allclose(hid[0], torch.from_numpy(g['x_h']), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('tag', ['surrogate', 'transplant'])
def test_bmcnet_oracle_vs_reference(golden_dir, tag, request):
    g = np.load(os.path.join(golden_dir, 'model_bmcnet_%s.npz' % tag))
    tr = request.getfixturevalue('plain_ckpt') if tag == 'transplant' else None
    sd = M.surrogate_state_dict(plain=False, seed=int(g['seed']), transplant=tr)
    outs, hid = _rollout(M.bmcnet_forward, sd, g, 4)
    assert torch.allclose(outs, torch.from_numpy(g['x_o']), atol=1e-5, rtol=1e-5)
    for t, k in zip(hid, ('x_h', 'x_h_p', 'x_h_n')):
        assert torch.allclose(t, torch.from_numpy(g[k]), atol=1e-5, rtol=1e-5), k


@pytest.mark.parametrize('plain', [False, True])
def test_state_dict_contract(golden_dir, plain):
    g = np.load(os.path.join(golden_dir, 'statedict_%s.npz' % ('plain' if plain else 'bmcnet')))
    keys = M.bmcnet_state_dict_keys(5, plain)
    assert keys == [str(k) for k in g['keys']]
    assert len(keys) == (120 if plain else 318)
    sd = M.surrogate_state_dict(plain=plain)
    assert [str(tuple(sd[k].shape)) for k in keys] == [str(s) for s in g['shapes']]
    ptr = {}
    groups = [ptr.setdefault(sd[k].data_ptr(), len(ptr)) for k in keys]
    assert groups == [int(v) for v in g['alias_group']]
    assert sum(sd[r].numel() for r in {M._alias_root(k) for k in keys}) == int(g['n_unique_params'])


# ---------------------------------------------------------------- dataloader window pipeline
def test_event_formatting_oracle_bit_exact_vs_reference(golden_dir):
    from oracle import h5windows_np as H
    d = np.load(os.path.join(golden_dir, 'fmt_events.npz'))
    for name in ('w2048', 'rel2048', 'w5', 'w1'):
        got = H.event_formatting((d[name + '_xs'], d[name + '_ys'], d[name + '_ts'], d[name + '_ps']))
        assert got.dtype == np.float32 and np.array_equal(got, d[name + '_out'], equal_nan=True), name


@pytest.mark.parametrize('n,window,sliding', [(10000, 2048, 1024), (5000, 1024, 512), (2049, 2048, 1024), (100, 2048, 1024), (4096, 1024, 0)])
def test_compute_k_indices_properties(n, window, sliding):
    """h5dataset.py:169-175,197-210: int(n / stride) windows, starts stride apart, ends clipped to n - 1."""
    from oracle import h5windows_np as H
    from bmcnet_esr_b200.dataloader.h5windows import compute_k_indices
    k = H.compute_k_indices(n, window, sliding)
    assert k == compute_k_indices(n, window, sliding)
    stride = window - sliding
    assert len(k) == n // stride
    for i, (a, b) in enumerate(k):
        assert a == i * stride and b == min(a + window, n - 1) and b < n
    assert H.compute_k_indices(n, window, sliding, dataset_length=1) == k[:1]
    assert H.compute_k_indices(n, window, sliding, dataset_length=10 ** 9) == k


# ---------------------------------------------------------------- inverse encoders
def test_redistribute_oracle_exact_vs_reference(golden_dir):
    from oracle import redistribute_ref as R
    d = np.load(os.path.join(golden_dir, 'redistribute.npz'))
    assert np.array_equal(R.event_redistribute(torch.from_numpy(d['pol_in']), True).numpy(), d['pol_out'])
    assert np.array_equal(R.event_redistribute(torch.from_numpy(d['nop_in']), False).numpy(), d['nop_out'])
    assert np.array_equal(R.event_redistribute(torch.zeros(2, 3, 4, 5), False).numpy(), d['zero_out'])
    assert np.array_equal(R.stack2cnt(torch.from_numpy(d['s2c_in'])).numpy(), d['s2c_out'])


def test_sequence_tuples_are_views_of_consecutive_windows():
    from bmcnet_esr_b200.dataloader.h5windows import sequence_item, sequence_tuples
    cnt = torch.arange(7 * 2 * 3 * 4, dtype=torch.float32).view(7, 2, 3, 4)
    t = sequence_tuples(cnt, 2)
    assert tuple(t.shape) == (6, 2, 2, 3, 4) and t.data_ptr() == cnt.data_ptr()
    for m in range(6):
        assert torch.equal(t[m], torch.stack([cnt[m], cnt[m + 1]]))        # concat_dict: torch.stack(dim=1) at batch 1
    item = sequence_item(cnt, 1, sequence_length=5, seqn=2, step_size=1)     # windows 1..5 -> 4 tuples
    assert len(item) == 4 and tuple(item[0].shape) == (1, 2, 2, 3, 4)
    assert torch.equal(item[0][0, 0], cnt[1]) and torch.equal(item[3][0, 1], cnt[5])
    assert torch.equal(item[0].transpose(1, 2)[0, :, 1], cnt[2])            # input_stack = inp_cnt.transpose(1, 2), infer_BMCNet.py:50
    with pytest.raises(IndexError):
        sequence_item(cnt, 3, sequence_length=5)


# ---- training iteration: oracle/train_step.py against the reference's own modules + Adam ----
import pytest as _pytest


@_pytest.mark.parametrize('tag', ['plain', 'bmcnet'])
def test_train_iteration_matches_reference(golden_dir, tag):
    """Two iterations of train.py:202-237 (sequence loss from zero state, one backward, Adam(amsgrad) with weight
    decay): losses, every unique parameter's gradient fingerprint after iteration 1 and value after iteration 2."""
    import numpy as np
    import torch
    from oracle import bmcnet_fp32 as O
    from oracle import train_step as T
    g = np.load(os.path.join(golden_dir, 'train_step_%s.npz' % tag))
    plain = tag == 'plain'
    sd = O.surrogate_state_dict(plain=plain, seed=2024)
    xs = [torch.from_numpy(a) for a in g['x']]
    gts = [torch.from_numpy(a) for a in g['gt']]

    def stats(t, k=24):
        f = t.detach().double().flatten()
        return np.concatenate([[f.sum().item(), f.abs().sum().item()], f[:k].numpy(), np.zeros(max(0, k - f.numel()))])

    state, losses = None, []
    for it in range(2):
        loss, grads, params, state = T.train_iteration(sd, xs, gts, plain, opt_state=state)
        losses.append(loss.item())
        if it == 0:
            keys = [k[5:] for k in g.files if k.startswith('grad.')]
            assert sorted(keys) == sorted(grads), 'unique-parameter set differs from the reference'
            for k in keys:
                want, got = g['grad.' + k], stats(grads[k])
                scale = max(1e-12, want[1] / max(1, grads[k].numel()))          # mean |grad|
                assert abs(got[0] - want[0]) <= 2e-3 * want[1] + 1e-9, (k, got[0], want[0])
                assert abs(got[1] - want[1]) <= 1e-3 * want[1] + 1e-9, (k, got[1], want[1])
                assert np.abs(got[2:] - want[2:]).max() <= 2e-2 * scale + 1e-3 * np.abs(want[2:]).max(), k
        sd = {k: params[O._alias_root(k)] for k in sd}              # next iteration sees the updated weights
    assert np.allclose(losses, g['losses'], rtol=1e-5), (losses, g['losses'])
    for k in [k[6:] for k in g.files if k.startswith('param.')]:
        want, got = g['param.' + k], stats(params[k])
        assert np.abs(got[2:] - want[2:]).max() <= 2e-6 + 1e-5 * np.abs(want[2:]).max(), k     # lr = 1e-4 steps
        assert abs(got[1] - want[1]) <= 1e-5 * want[1] + 1e-9, k


def test_adam_amsgrad_restatement_matches_torch():
    import torch
    from oracle import train_step as T
    torch.manual_seed(0)
    p0 = {'a': torch.randn(7, 5), 'b': torch.randn(11)}
    ref = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    opt = torch.optim.Adam(ref.values(), lr=1e-3, weight_decay=1e-2, amsgrad=True)
    mine, state = {k: v.clone() for k, v in p0.items()}, {}
    for step in range(5):
        grads = {k: torch.randn_like(v) * (0.1 if step == 3 else 1.0) for k, v in p0.items()}      # a small step exercises vmax
        for k in ref:
            ref[k].grad = grads[k].clone()
        opt.step()
        T.adam_amsgrad_step(mine, grads, state, lr=1e-3, weight_decay=1e-2)
    for k in ref:
        assert torch.allclose(mine[k], ref[k].detach(), rtol=1e-6, atol=1e-7), k
