"""GPU parity of the inverse encoders (csrc/redistribute.cu through bmcnet_esr_b200/dataloader/encodings.py) against
the reference-generated golden and the CPU oracle (oracle/redistribute_ref.py).

Bars: coordinates, polarities, lengths, padding and stack2cnt bit-exact; timestamps within 2e-7 absolute (they are
torch.linspace values in (0, 1]: ATen's vectorised kernel evaluates start + step * k in chunks whose boundaries
depend on the CPU's vector width, so the last bit is not defined by the reference's source); per entry the events
are non-decreasing in t and, as a multiset, equal to the reference's."""
import os

import numpy as np
import pytest
import torch

from oracle import redistribute_ref as O
from oracle.make_golden import synth_stack

pytestmark = pytest.mark.gpu
T_TOL = 2e-7


def _compare(got, ref):
    got, ref = got.cpu().numpy(), np.asarray(ref)
    assert got.shape == ref.shape
    for b in range(ref.shape[0]):
        n = int((ref[b, :, 3] != 0).sum())
        assert not got[b, n:].any() and not ref[b, n:].any()                       # zero padding
        g, r = got[b, :n], ref[b, :n]
        assert np.all(np.diff(g[:, 2]) >= 0), 'entry %d not sorted by t' % b
        assert np.abs(np.sort(g[:, 2]) - np.sort(r[:, 2])).max(initial=0) <= T_TOL
        # same multiset of events: timestamps that differ by an ulp may swap neighbours, nothing else
        kg = np.lexsort((g[:, 3], g[:, 1], g[:, 0], np.round(g[:, 2] * 1e5)))
        kr = np.lexsort((r[:, 3], r[:, 1], r[:, 0], np.round(r[:, 2] * 1e5)))
        assert np.array_equal(g[kg][:, [0, 1, 3]], r[kr][:, [0, 1, 3]])
        assert np.abs(g[kg][:, 2] - r[kr][:, 2]).max(initial=0) <= 1e-5 + T_TOL


def test_against_reference_golden(golden_dir):
    from bmcnet_esr_b200.dataloader import encodings as G
    d = np.load(os.path.join(golden_dir, 'redistribute.npz'))
    _compare(G.python_event_redistribute_PolarityStack(torch.from_numpy(d['pol_in']).cuda()), d['pol_out'])
    _compare(G.python_event_redistribute_NoPolarityStack(torch.from_numpy(d['nop_in']).cuda()), d['nop_out'])
    z = G.python_event_redistribute_NoPolarityStack(torch.zeros(2, 3, 4, 5).cuda())
    assert tuple(z.shape) == (2, 1, 4) and not z.any()
    assert np.array_equal(G.stack2cnt(torch.from_numpy(d['s2c_in']).cuda()).cpu().numpy(), d['s2c_out'])


@pytest.mark.parametrize('b,shape,polarity,seed', [(2, (2, 5, 9, 11), True, 1), (4, (5, 12, 16), False, 2), (1, (1, 3, 3), False, 3)])
def test_against_oracle_random(b, shape, polarity, seed):
    from bmcnet_esr_b200.dataloader import encodings as G
    st = synth_stack(b, shape, 50 + seed, rate=0.2, vmax=6)
    if polarity:
        st = st.abs()
    fn = G.python_event_redistribute_PolarityStack if polarity else G.python_event_redistribute_NoPolarityStack
    _compare(fn(st.cuda()), O.event_redistribute(st, polarity).numpy())
    if not polarity:
        assert np.array_equal(G.stack2cnt(st.cuda()).cpu().numpy(), O.stack2cnt(st).numpy())


def test_round_trip_through_time_binning():
    """stack -> events -> per-bin signed counts gives the stack back (the round trip the reference's own
    `__main__` attempts, encodings.py:674-697): linspace timestamps stay inside their bin."""
    from bmcnet_esr_b200.dataloader import encodings as G
    B, C, H, W = 1, 5, 20, 24
    st = synth_stack(B, (C, H, W), 77, rate=0.3, vmax=3).round()
    ev = G.python_event_redistribute_NoPolarityStack(st.cuda())[0]
    n = int((ev[:, 3] != 0).sum())
    xs, ys, ts, ps = (ev[:n, k].contiguous() for k in range(4))
    # timestamps of bin c lie in (c/C, (c+1)/C]: the last event of a voxel sits exactly on the upper edge
    cnt = torch.zeros(C, H, W, device='cuda')
    bins = torch.clamp(torch.ceil(ts * C - 1e-4).long() - 1, min=0, max=C - 1)
    cnt.index_put_((bins, ys.long(), xs.long()), ps, accumulate=True)
    assert torch.equal(cnt.cpu(), st[0])
    assert float(ts.min()) > 0 and float(ts.max()) <= 1
    # large sparse stack: lengths and sortedness only (the sort runs one CTA per entry)
    big = synth_stack(3, (5, 90, 160), 78, rate=0.05, vmax=5)
    out = G.python_event_redistribute_NoPolarityStack(big.cuda())
    tot = big.round().abs().sum(dim=(1, 2, 3))
    assert out.shape[1] == int(tot.max())
    for b in range(3):
        nb = int(tot[b])
        assert bool((out[b, 1:nb, 2] >= out[b, :nb - 1, 2]).all()) and not bool(out[b, nb:].any())


def test_random_mode_stays_inside_the_bins():
    from bmcnet_esr_b200.dataloader import encodings as G
    st = synth_stack(2, (4, 8, 8), 79, rate=0.3, vmax=3)
    ev = G.python_event_redistribute_NoPolarityStack(st.cuda(), mode='random')
    lin = G.python_event_redistribute_NoPolarityStack(st.cuda(), mode='linear')
    assert ev.shape == lin.shape
    for b in range(2):
        n = int((lin[b, :, 3] != 0).sum())
        assert bool((ev[b, 1:n, 2] >= ev[b, :n - 1, 2]).all())
        bin_r = torch.clamp(torch.ceil(ev[b, :n, 2] * 4 - 1e-4).long() - 1, min=0, max=3)
        bin_l = torch.clamp(torch.ceil(lin[b, :n, 2] * 4 - 1e-4).long() - 1, min=0, max=3)
        assert torch.equal(bin_r.sort()[0], bin_l.sort()[0])
