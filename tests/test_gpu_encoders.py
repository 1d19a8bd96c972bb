"""GPU parity of the event encoders (csrc/encode.cu through the drop-in functions of
bmcnet_esr_b200/dataloader/encodings.py) against the CPU oracle and the reference goldens.

Bars: count-valued encodings bit-exact, incl. the reference's in-place side
effects and quirks; float-weighted voxels within 1e-6 relative (relative to the
largest voxel: the reference sums serially in fp32, the GPU in another order)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import encodings_np as E
from oracle.make_golden import synth_events

pytestmark = pytest.mark.gpu
VOXEL_RTOL = 1e-6      # north_star bar for float-weighted encodings, relative to max|ref|


@pytest.fixture(scope='module')
def G():
    from bmcnet_esr_b200.dataloader import encodings
    return encodings


def _gpu(ev):
    return [torch.from_numpy(a.copy()).cuda() for a in ev]


EXACT = {
    'channels': lambda M, a, h, w, B: M.events_to_channels(a[0], a[1], a[3], sensor_size=(h, w)),
    'stack_polarity': lambda M, a, h, w, B: M.events_to_stack_polarity(a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'stack_no_polarity': lambda M, a, h, w, B: M.events_to_stack_no_polarity(a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'voxel_torch_hard': lambda M, a, h, w, B: M.events_to_voxel_torch(a[0], a[1], a[2], a[3], B, sensor_size=(h, w),
                                                                      temporal_bilinear=False),
    'image': lambda M, a, h, w, B: M.events_to_image(a[0], a[1], a[3], sensor_size=(h, w)),
    'image_torch': lambda M, a, h, w, B: M.events_to_image_torch(a[0], a[1], a[3], sensor_size=(h, w)),
}
FLOAT = {
    'voxel': lambda M, a, h, w, B: M.events_to_voxel(a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'voxel_torch': lambda M, a, h, w, B: M.events_to_voxel_torch(a[0], a[1], a[2], a[3], B, sensor_size=(h, w)),
    'image_torch_bilinear': lambda M, a, h, w, B: M.events_to_image_torch(a[0], a[1], a[3], sensor_size=(h, w),
                                                                          interpolation='bilinear'),
}


def _float_close(got, ref, ev, h, w):
    """1e-6 relative (to the largest value) while no output pixel sums more than 16 events --
    the regime of the reference's own windows.  The reference adds serially in fp32, the GPU in
    another order: with K events on one pixel both carry up to K half-ulps of rounding, so the
    bar widens to 2^-24 * K for heavier pixels (e.g. the (0,0) pixel that collects every
    out-of-range event)."""
    xs, ys = ev[0], ev[1]
    oor = (xs >= w) | (xs < 0) | (ys >= h) | (ys < 0)
    pix = np.where(oor, 0, np.trunc(ys).astype(np.int64) * w + np.trunc(xs).astype(np.int64))
    kmax = int(np.bincount(pix, minlength=1).max()) if len(pix) else 0
    tol = max(VOXEL_RTOL, 2.0 ** -24 * kmax) * max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(got - ref).max())
    return err <= tol, err, tol


@pytest.mark.parametrize('fname', sorted(EXACT) + sorted(FLOAT))
def test_against_reference_goldens(G, golden_dir, fname):
    for path in sorted(glob.glob(os.path.join(golden_dir, 'enc_*.npz'))):
        g = np.load(path)
        h, w, B = int(g['h']), int(g['w']), int(g['B'])
        a = _gpu([g['in_' + k] for k in ('xs', 'ys', 'ts', 'ps')])
        got = (EXACT.get(fname) or FLOAT[fname])(G, a, h, w, B).cpu().numpy()
        ref = g['out_' + fname]
        assert got.shape == ref.shape, (path, fname)
        if fname in EXACT:
            assert np.array_equal(got, ref), (path, fname, float(np.abs(got - ref).max()))
        else:
            ok, err, tol = _float_close(got, ref, [g['in_' + k] for k in ('xs', 'ys', 'ts', 'ps')], h, w)
            assert ok, (path, fname, err, tol)
        for i, k in enumerate(('xs', 'ys', 'ts', 'ps')):          # in-place side effects
            key = 'mut_%s_%s' % (fname, k)
            want = g[key] if key in g.files else g['in_' + k]
            assert np.array_equal(a[i].cpu().numpy(), want), (path, fname, k)


@pytest.mark.parametrize('n,h,w,B', [(1, 5, 7, 3), (5, 5, 7, 2), (4097, 45, 80, 5), (200000, 45, 80, 5),
                                     (150000, 180, 320, 3), (450000, 180, 320, 5), (100000, 360, 640, 2), (70000, 31, 56, 9)])
def test_against_oracle_random(G, n, h, w, B):
    """Seeded random streams incl. out-of-range, fractional and duplicate-timestamp events; every
    histogram tier (smem int32 / packed 16-bit / global / one time bin per CTA) is hit by one of the sizes."""
    ev = synth_events(n, h, w, seed=n % 97, oor=0.04, dup=True, frac=True)
    for name, fn in {**EXACT, **FLOAT}.items():
        ca, ga = [x.copy() for x in ev], _gpu(ev)
        ref = fn(E, ca, h, w, B)
        got = fn(G, ga, h, w, B).cpu().numpy()
        assert got.shape == ref.shape, name
        if name in EXACT:
            assert np.array_equal(got, ref), (name, float(np.abs(got - ref).max()))
        else:
            ok, err, tol = _float_close(got, ref, ev, h, w)
            assert ok, (name, err, tol)
        for c, gt in zip(ca, ga):
            assert np.array_equal(c, gt.cpu().numpy()), name


def test_empty_and_unaligned_inputs(G):
    z = torch.zeros(0, device='cuda')
    out = G.events_to_channels(z, z.clone(), z.clone(), sensor_size=(6, 9))
    assert out.shape == (2, 6, 9) and float(out.abs().sum()) == 0
    # a sliced (4-byte aligned only) view must take the scalar path and still be exact
    ev = synth_events(10001, 12, 16, seed=3, oor=0.05, frac=True)
    full = _gpu(ev)
    view = [t[1:].contiguous() for t in full]
    shifted = [torch.cat([t[:1], t[1:]])[1:] for t in full]        # storage offset 1 -> unaligned
    assert shifted[0].data_ptr() % 16 != 0
    a = G.events_to_channels(view[0], view[1], view[3], sensor_size=(12, 16))
    b = G.events_to_channels(shifted[0], shifted[1], shifted[3], sensor_size=(12, 16))
    assert torch.equal(a, b)


def test_windows_match_per_window_calls(G):
    h, w, win = 45, 80, 2048
    ev = synth_events(win * 37 + 111, h, w, seed=5, oor=0.02, frac=True)
    a = _gpu(ev)
    n = len(ev[0])
    offs = torch.tensor(list(range(0, n, win)) + [n], dtype=torch.int64, device='cuda')
    got = G.events_to_channels_windows(a[0], a[1], a[3], offs, sensor_size=(h, w))
    b = _gpu(ev)
    for i in range(len(offs) - 1):
        s, e = int(offs[i]), int(offs[i + 1])
        ref = E.events_to_channels(ev[0][s:e].copy(), ev[1][s:e].copy(), ev[3][s:e].copy(), sensor_size=(h, w))
        assert np.array_equal(got[i].cpu().numpy(), ref), i
    del b


def test_full_size_properties(G):
    """BASELINE sizes (1e8 events; 1e9 needs 12 GB and is exercised by bench.py): properties that do
    not need the CPU oracle -- total count, additivity over a split of the stream, order invariance."""
    n, h, w = 100_000_000, 45, 80
    g = torch.Generator(device='cuda').manual_seed(1)
    xs = torch.rand(n, device='cuda', generator=g) * w
    ys = torch.rand(n, device='cuda', generator=g) * h
    ps = (torch.rand(n, device='cuda', generator=g) < 0.5).float() * 2 - 1
    full = G.events_to_channels(xs, ys, ps, sensor_size=(h, w))
    assert float(full.sum(dtype=torch.float64)) == n
    assert float(full[0].sum(dtype=torch.float64)) == float((ps > 0).sum())
    k = 33_333_337
    parts = G.events_to_channels(xs[:k].contiguous(), ys[:k].contiguous(), ps[:k].contiguous(), sensor_size=(h, w)) + \
        G.events_to_channels(xs[k:].contiguous(), ys[k:].contiguous(), ps[k:].contiguous(), sensor_size=(h, w))
    assert torch.equal(full, parts)
    perm = torch.randperm(n, device='cuda', generator=g)
    assert torch.equal(full, G.events_to_channels(xs[perm], ys[perm], ps[perm], sensor_size=(h, w)))
    # oracle on a 2e6 prefix, and the HR grid through the 16-bit packed tier with a hot pixel
    m = 2_000_000
    ref = E.events_to_channels(xs[:m].cpu().numpy(), ys[:m].cpu().numpy(), ps[:m].cpu().numpy(), sensor_size=(h, w))
    assert np.array_equal(G.events_to_channels(xs[:m].contiguous(), ys[:m].contiguous(), ps[:m].contiguous(),
                                               sensor_size=(h, w)).cpu().numpy(), ref)
    xh = xs[:20_000_000] * 4
    yh = ys[:20_000_000] * 4
    xh[:5_000_000] = 17.0          # 5e6 events on one pixel: forces the 16-bit carry path many times
    yh[:5_000_000] = 3.0
    ph = torch.ones_like(xh)
    hr = G.events_to_channels(xh, yh, ph, sensor_size=(180, 320))
    assert float(hr.sum(dtype=torch.float64)) == 20_000_000
    assert float(hr[0, 180 - 1 - 3, 17]) >= 5_000_000
    cnt = torch.zeros(180 * 320, device='cuda', dtype=torch.float64)
    cnt.index_add_(0, ((180 - 1 - yh.long()) * 320 + xh.long()), torch.ones_like(xh, dtype=torch.float64))
    assert torch.equal(hr[0].double().flatten(), cnt)


def test_fp32_saturation_matches_reference_semantics(G):
    # the reference adds 1.0f serially and sticks at 2^24; 2^24 + 1000 events on one pixel
    n = (1 << 24) + 1000
    xs = torch.full((n,), 2.0, device='cuda')
    ys = torch.full((n,), 1.0, device='cuda')
    ps = torch.ones(n, device='cuda')
    out = G.events_to_channels(xs, ys, ps, sensor_size=(4, 4))
    assert float(out[0, 4 - 1 - 1, 2]) == float(1 << 24)


def test_full_size_voxel_and_stack_properties(G):
    """BASELINE-size streams (1e8 events) through the float and the binned encoders: properties that need no
    CPU oracle -- mass conservation, additivity over a split of the stream, the y-flip identity between the two
    voxel entry points, and the bin structure of the stack encoders."""
    n, h, w, B = 100_000_000, 45, 80, 5
    g = torch.Generator(device='cuda').manual_seed(2)
    xs = torch.floor(torch.rand(n, device='cuda', generator=g) * w)
    ys = torch.floor(torch.rand(n, device='cuda', generator=g) * h)
    ps = (torch.rand(n, device='cuda', generator=g) < 0.5).float() * 2 - 1
    ts = torch.sort(torch.rand(n, device='cuda', generator=g))[0]
    ts = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)
    vox = G.events_to_voxel(xs, ys, ts, ps, B, sensor_size=(h, w))
    # the two temporal weights of an event sum to 1: total mass = sum(p) (fp32 bins, fp64 reduction)
    assert abs(float(vox.sum(dtype=torch.float64)) - float(ps.sum(dtype=torch.float64))) <= 1e-6 * n
    k = 41_234_567
    parts = G.events_to_voxel(xs[:k].contiguous(), ys[:k].contiguous(), ts[:k].contiguous(), ps[:k].contiguous(), B, sensor_size=(h, w)) + \
        G.events_to_voxel(xs[k:].contiguous(), ys[k:].contiguous(), ts[k:].contiguous(), ps[k:].contiguous(), B, sensor_size=(h, w))
    scale = float(vox.abs().max())
    assert float((vox - parts).abs().max()) <= 2e-5 * scale           # ~5500 fp32 additions per bin in another order
    # events_to_voxel_torch == flip(events_to_voxel) on pre-normalised ts
    vt = G.events_to_voxel_torch(xs, ys, ts, ps, B, sensor_size=(h, w))
    assert float((vt - torch.flip(vox, dims=[1])).abs().max()) <= 2e-5 * scale
    # large grid: the pair-reduction path agrees with the shared-memory path on the same events (coordinates x1)
    big = G.events_to_voxel(xs[:20_000_000].contiguous(), ys[:20_000_000].contiguous(), ts[:20_000_000].contiguous(),
                            ps[:20_000_000].contiguous(), B, sensor_size=(180, 320))
    small = G.events_to_voxel(xs[:20_000_000].contiguous(), ys[:20_000_000].contiguous(), ts[:20_000_000].contiguous(),
                              ps[:20_000_000].contiguous(), B, sensor_size=(h, w))
    # y-flip: row H-1-y of the small grid is row 179-y of the large one; columns coincide
    assert float((big[:, 180 - h:, :w] - small).abs().max()) <= 2e-5 * float(small.abs().max())
    assert float(big[:, :180 - h].abs().max()) == 0.0 and float(big[:, :, w:].abs().max()) == 0.0
    # stack encoders: every event lands in >= 1 bin, boundary events in two; polarity planes are counts
    st = G.events_to_stack_polarity(xs, ys, ts, ps, B, sensor_size=(h, w))
    total = float(st.sum(dtype=torch.float64))
    assert n <= total <= n + 2 * B
    assert float(st.min()) >= 0.0
    sn = G.events_to_stack_no_polarity(xs, ys, ts, ps, B, sensor_size=(h, w))
    assert torch.equal(sn, st[0] - st[1])
    # bins partition time: bin b of the stack equals the count image of the events in its time slice
    edges = torch.searchsorted(ts, torch.linspace(0, 1, B + 1, device='cuda')[1:-1].contiguous())
    lo, hi = int(edges[1]) + 8, int(edges[2]) - 8                     # strictly inside bin 2
    mid = G.events_to_image_torch(xs[lo:hi].contiguous(), ys[lo:hi].contiguous(), ps[lo:hi].contiguous(), sensor_size=(h, w))
    assert float((sn[2] - mid).abs().max()) <= 16.0                   # up to 8 events on either side of the slice


def test_stack_one_bin_per_cta_kernel_unaligned_and_sharded(G):
    """stack_bins_kernel (grids whose planes fit shared memory only one time bin at a time): a 4-byte-aligned view
    (scalar path), and two event ranges through the shard entry, against the oracle on the whole stream."""
    from bmcnet_esr_b200 import sharding as S
    n, h, w, B = 500_001, 180, 320, 5
    ev = synth_events(n, h, w, seed=21, oor=0.03, dup=True, frac=True)
    for pol, name in ((True, 'stack_polarity'), (False, 'stack_no_polarity')):
        ref = EXACT[name](E, [x[1:].copy() for x in ev], h, w, B)
        odd = [t[1:] for t in _gpu(ev)]                            # contiguous views with data_ptr % 16 == 4
        assert odd[0].data_ptr() % 16 == 4
        got = EXACT[name](G, odd, h, w, B).cpu().numpy()
        assert np.array_equal(got, ref), name
        # two ranks' worth of events, summed by hand
        full = [x[1:].copy() for x in ev]
        ts_all = torch.from_numpy(full[2]).cuda()
        acc = None
        for lo, hi in ((0, 233_332), (233_332, len(full[0]))):
            cut = [torch.from_numpy(full[i][lo:hi].copy()).cuda() for i in (0, 1, 3)]
            fn = S.events_to_stack_polarity_sharded if pol else S.events_to_stack_no_polarity_sharded
            part = fn(cut[0], cut[1], ts_all, cut[2], lo, B, sensor_size=(h, w))
            acc = part if acc is None else acc + part
        assert np.array_equal(acc.cpu().numpy(), ref), name + ' sharded'


@pytest.mark.parametrize('h,w,B,n', [(45, 80, 5, 3_000_017), (180, 320, 5, 2_000_003), (22, 40, 1, 70_001)])
def test_deterministic_float_path_is_bit_reproducible_and_exact(G, h, w, B, n):
    """BMC_ENC_DETERMINISTIC (`with deterministic():`): the float-weighted encodings accumulate in 64-bit fixed
    point.  (1) two runs give identical bits -- also against a run on a shifted (unaligned -> scalar loads, another
    grid) copy of the same events, i.e. a different summation order; (2) the value is the correctly rounded exact sum
    up to 2^-33 per event: compared with a float64 accumulation of the SAME fp32 weights it agrees to 1 ulp of fp32;
    (3) the default atomic path stays within the 1e-6 bar of it."""
    ev = synth_events(n, h, w, seed=9, oor=0.0)

    def run(det, shift=False):
        a = _gpu(ev)
        if shift:                                   # same values at a 4-byte offset: the kernel takes its scalar path
            a = [torch.cat([t.new_zeros(1), t])[1:] for t in a]
            assert a[0].data_ptr() % 16 != 0
            a = [t for t in a]
        with G.deterministic(det):
            # views of a larger tensor are still contiguous 1-D float32
            return G.events_to_voxel(a[0], a[1], a[2], a[3], B, sensor_size=(h, w)).cpu().numpy()

    d1, d2, d3 = run(True), run(True), run(True, shift=True)
    assert np.array_equal(d1, d2) and np.array_equal(d1, d3)
    # float64 accumulation of the same fp32 weights (the oracle's weights, summed without fp32 rounding)
    xs, ys, ts, ps = ev
    tn = (ts * np.float32(B - 1)).astype(np.float32)
    ref = np.zeros((B, h, w), np.float64)
    yy = (h - 1 - np.trunc(ys).astype(np.int64))
    xx = np.trunc(xs).astype(np.int64)
    for b in range(B):
        wgt = np.maximum(np.float32(0), np.float32(1) - np.abs(tn - np.float32(b))).astype(np.float32)
        np.add.at(ref[b], (yy, xx), (ps * wgt).astype(np.float32).astype(np.float64))
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(d1 - ref).max() <= 2.0 ** -23 * scale, np.abs(d1 - ref).max()
    nd = run(False)
    assert np.abs(nd - d1).max() <= max(VOXEL_RTOL, 2.0 ** -24 * (n / (h * w)) * 4) * scale
    if 2 * B * h * w * 4 <= 227 * 1024:
        # grids that fit shared memory twice: the DEFAULT path is the 2^-24 fixed-point one (integer adds only), hence
        # bit-reproducible as well, and within 2^-25 per event of the exact sum
        assert np.array_equal(nd, run(False)) and np.array_equal(nd, run(False, shift=True))
        assert np.abs(nd - ref).max() <= (2.0 ** -23 + 2.0 ** -25 * 4) * scale * max(1.0, n / (h * w * B) / 8)


def test_deterministic_image_paths(G):
    h, w, n = 45, 80, 400_003
    ev = list(synth_events(n, h, w, seed=4, oor=0.01))
    rng = np.random.default_rng(1)
    ev[3] = (ev[3] * rng.random(n).astype(np.float32)).astype(np.float32)          # non-unit weights -> float path

    def run(fn, **kw):
        outs = []
        for _ in range(2):
            a = _gpu(ev)
            with G.deterministic():
                outs.append(fn(a[0], a[1], a[3], sensor_size=(h, w), **kw).cpu().numpy())
        assert np.array_equal(outs[0], outs[1])
        a = _gpu(ev)
        base = fn(a[0], a[1], a[3], sensor_size=(h, w), **kw).cpu().numpy()
        assert np.abs(base - outs[0]).max() <= 1e-5 * max(1.0, np.abs(base).max())
    run(G.events_to_image)
    run(G.events_to_image_torch)
    run(G.events_to_image_torch, interpolation='bilinear')


# ---------------------------------------------------------------------------------------------- bins split over CTAs (role_kernel)
class _cluster_bins:
    def __init__(self, G, on=True):
        self.G, self.on = G, on

    def __enter__(self):
        self.prev, self.G.SPLIT_BINS = self.G.SPLIT_BINS, self.on

    def __exit__(self, *exc):
        self.G.SPLIT_BINS = self.prev


@pytest.mark.parametrize('n,h,w,B,sort_t', [(300_017, 180, 320, 5, True), (450_000, 180, 320, 3, False), (250_001, 360, 640, 2, True),
                                            (5, 360, 640, 2, True), (120_000, 120, 300, 9, True), (400_003, 360, 640, 3, True)])
def test_cluster_bins_against_oracle(G, n, h, w, B, sort_t):
    """role_kernel (a few CTAs stream the same events, each holding a share of the bins in shared memory) forced at small
    event counts: events_to_channels bit-exact incl. the F9 leak and the (deferred) in-place zeroing, events_to_voxel /
    events_to_voxel_torch within the 1e-6 bar -- on time-sorted input (the fast path) and on shuffled timestamps (most
    events take the global-atomic path)."""
    ev = list(synth_events(n, h, w, seed=n % 89, oor=0.04, dup=True, frac=True))
    if not sort_t:
        ev[2] = np.random.default_rng(3).permutation(ev[2])
    cases = {'channels': EXACT['channels'], 'voxel': FLOAT['voxel'], 'voxel_torch': FLOAT['voxel_torch']}
    for name, fn in cases.items():
        if name == 'voxel_torch' and not sort_t:
            continue                                  # the reference's ts[0] / ts[-1] normalisation assumes sorted stamps
        ca, ga = [x.copy() for x in ev], _gpu(ev)
        ref = fn(E, ca, h, w, B)
        with _cluster_bins(G):
            got = fn(G, ga, h, w, B).cpu().numpy()
        assert got.shape == ref.shape, name
        if name == 'channels':
            assert np.array_equal(got, ref), (name, float(np.abs(got - ref).max()))
        else:
            ok, err, tol = _float_close(got, ref, ev, h, w)
            assert ok, (name, err, tol)
        for c, gt in zip(ca, ga):
            assert np.array_equal(c, gt.cpu().numpy()), name


def test_cluster_bins_skewed_and_heavy_pixels(G):
    """Adversarial streams: (1) every event on ONE pixel -- the 16-bit counters / 32-bit fixed-point sums carry many
    times; (2) all events in one role's share of the pixels; (3) more out-of-range events than the deferred-zeroing list
    holds.  Counts stay bit-exact against bincount, voxels are bit-reproducible and match a float64 sum."""
    h, w, n = 360, 640, 3_000_000
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(11)
    ps = (torch.randint(0, 2, (n,), device=dev, generator=g) * 2 - 1).float()
    for mode in ('one_pixel', 'one_tile'):
        if mode == 'one_pixel':
            xs, ys = torch.full((n,), 17.0, device=dev), torch.full((n,), 300.0, device=dev)
        else:
            xs = torch.randint(0, w, (n,), device=dev, generator=g).float()
            ys = torch.randint(0, 40, (n,), device=dev, generator=g).float()
        with _cluster_bins(G):
            got = G.events_to_channels(xs.clone(), ys.clone(), ps.clone(), sensor_size=(h, w))
        pix = ((h - 1 - ys.long()) * w + xs.long()) + (ps < 0).long() * h * w
        ref = torch.bincount(pix, minlength=2 * h * w).view(2, h, w).float()
        assert torch.equal(got, ref), mode
    # (3) 200k out-of-range events (> 65536 list entries): the zeroing falls back to a scan; F9 leak at [1, H-1, 0]
    xs = torch.randint(0, w, (n,), device=dev, generator=g).float()
    ys = torch.randint(0, h, (n,), device=dev, generator=g).float()
    xs[:200_000] = float(w) + 3.0
    xc, yc = xs.clone(), ys.clone()
    with _cluster_bins(G):
        got = G.events_to_channels(xc, yc, ps, sensor_size=(h, w))
    inr = xs < w
    pix = ((h - 1 - ys[inr].long()) * w + xs[inr].long()) + (ps[inr] < 0).long() * h * w
    ref = torch.bincount(pix, minlength=2 * h * w).view(2, h, w).float()
    ref[1, h - 1, 0] += float((ps[:200_000] < 0).sum())
    assert torch.equal(got, ref)
    assert float(xc[:200_000].abs().sum()) == 0.0 and float(yc[:200_000].abs().sum()) == 0.0 and torch.equal(xc[200_000:], xs[200_000:])
    # voxels, 180x320: the forced split-bins path across two runs and against a float64 sum
    h, w, B = 180, 320, 5
    ts = torch.sort(torch.rand(n, device=dev, generator=g))[0]
    xs, ys = torch.full((n,), 5.0, device=dev), torch.full((n,), 7.0, device=dev)
    with _cluster_bins(G):
        a = G.events_to_voxel(xs.clone(), ys.clone(), ts, ps, B, sensor_size=(h, w))
        b = G.events_to_voxel(xs.clone(), ys.clone(), ts, ps, B, sensor_size=(h, w))
    assert torch.equal(a, b)                                       # integer sums: bit-reproducible
    tn = ts.double() * (B - 1)
    ref = torch.zeros(B, dtype=torch.float64, device=dev)
    for k in range(B):
        ref[k] = (ps.double() * (1 - (tn - k).abs()).clamp(min=0)).sum()
    got = a[:, h - 1 - 7, 5].double()
    assert float((got - ref).abs().max()) <= 2.0 ** -24 * n * 0.5 + 1e-3, (got, ref)
    assert float(a.abs().sum() - a[:, h - 1 - 7, 5].abs().sum()) == 0.0


def test_cluster_bins_full_size(G):
    """2^26 events (the default dispatch: no flag): 360x640 counts bit-exact against bincount; 180x320 voxels: total
    mass, and equality with the sum of the two halves of the stream encoded separately to 1e-6."""
    n, dev = 1 << 26, 'cuda'
    g = torch.Generator(device=dev).manual_seed(5)
    h, w = 360, 640
    xs = torch.randint(0, w, (n,), device=dev, generator=g).float()
    ys = torch.randint(0, h, (n,), device=dev, generator=g).float()
    ps = (torch.randint(0, 2, (n,), device=dev, generator=g) * 2 - 1).float()
    got = G.events_to_channels(xs, ys, ps, sensor_size=(h, w))
    pix = ((h - 1 - ys.long()) * w + xs.long()) + (ps < 0).long() * h * w
    assert torch.equal(got, torch.bincount(pix, minlength=2 * h * w).view(2, h, w).float())
    del got, pix
    h, w, B = 180, 320, 5
    xs, ys = (xs * 0.5).floor(), (ys * 0.5).floor()
    ts = torch.sort(torch.rand(n, device=dev, generator=g))[0]
    full = G.events_to_voxel(xs, ys, ts, ps, B, sensor_size=(h, w))
    assert abs(float(full.double().sum()) - float(ps.double().sum())) <= 1e-6 * n
    half = n // 2
    parts = sum(G.events_to_voxel(xs[s], ys[s], ts[s], ps[s], B, sensor_size=(h, w)) for s in (slice(0, half), slice(half, n)))
    assert float((full - parts).abs().max()) <= 1e-6 * max(1.0, float(full.abs().max())) * 4


def test_cluster_bins_unaligned_views(G):
    """The split-bins kernel on 4-byte-aligned views (scalar loads, single-event units) and on event counts that are not a
    multiple of 4: same bits as the aligned call."""
    n, dev = 700_001, 'cuda'
    g = torch.Generator(device=dev).manual_seed(21)
    ts = torch.sort(torch.rand(n + 1, device=dev, generator=g))[0]
    ps = (torch.randint(0, 2, (n + 1,), device=dev, generator=g) * 2 - 1).float()
    for (h, w), fn in (((180, 320), lambda x, y, t, p, hw: G.events_to_voxel(x, y, t, p, 5, sensor_size=hw)),
                       ((360, 640), lambda x, y, t, p, hw: G.events_to_channels(x, y, p, sensor_size=hw))):
        xs = torch.randint(-2, w + 2, (n + 1,), device=dev, generator=g).float()
        ys = torch.randint(0, h, (n + 1,), device=dev, generator=g).float()
        with _cluster_bins(G):
            a = fn(xs[1:].clone(), ys[1:].clone(), ts[1:].clone(), ps[1:].clone(), (h, w))           # aligned copies
            views = [t[1:] for t in (xs.clone(), ys.clone(), ts.clone(), ps.clone())]                   # storage offset 1
            assert views[0].data_ptr() % 16 != 0
            b = fn(*views, (h, w))
        assert torch.equal(a, b)
        with _cluster_bins(G, False):
            c = fn(xs[1:].clone(), ys[1:].clone(), ts[1:].clone(), ps[1:].clone(), (h, w))           # default path
        assert float((a - c).abs().max()) <= 1e-6 * max(1.0, float(c.abs().max())) * 4
