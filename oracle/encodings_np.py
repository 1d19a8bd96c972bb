"""CPU oracle for the event encoders -- TEST INFRASTRUCTURE, NOT PRODUCT.

numpy restatement of the forward encoders of the reference
(`dataloader/encodings.py:6-305`).  Only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this module;
the product (`bmcnet_esr_b200/`) never does.

Parity pin: the reference ships NO golden vectors or tests for this path
(SURVEY.md section 4), so this restatement is pinned against outputs of the
reference itself, generated in the build container by `oracle/make_golden.py`
(which imports `/root/reference`) and committed under `tests/golden/`.
`tests/test_oracle_vs_golden.py` re-checks every function below against them.

Everything is float32 with the reference's operation ORDER (two roundings where
the reference has two tensor ops, true division, python scalars cast to float32),
and every in-place side effect of the reference on its arguments is reproduced,
because later calls observe it: arrays passed in ARE mutated.

Arrays are 1-D float32 numpy arrays (the reference's input contract,
`dataloader/base_dataset.py:24-31`).
"""
import numpy as np

F32 = np.float32


def _f32(a):
    a = np.asarray(a)
    assert a.dtype == np.float32, "oracle expects float32 event arrays"
    return a


def _trunc_long(a):
    # torch `.long()` on float32 truncates toward zero (encodings.py:260-263, 67-70)
    return np.trunc(a).astype(np.int64)


def _oor_zero_inplace(xs, ys, ps, sensor_size):
    """encodings.py:249-254 and :34-39 -- out-of-range events are zeroed IN PLACE."""
    H, W = sensor_size
    mask = (xs >= W) | (xs < 0) | (ys >= H) | (ys < 0)
    xs[mask] = 0
    ys[mask] = 0
    ps[mask] = 0
    return mask


def events_to_image(xs, ys, ps, sensor_size=(180, 240)):
    """encodings.py:241-269.  y-flipped serial scatter-add; mutates xs, ys, ps."""
    xs, ys, ps = _f32(xs), _f32(ys), _f32(ps)
    H, W = sensor_size
    _oor_zero_inplace(xs, ys, ps, sensor_size)
    img = np.zeros((H, W), dtype=np.float32)
    xi = _trunc_long(xs)
    yi = H - _trunc_long(ys) - 1                     # :265 vertical flip
    np.add.at(img, (yi, xi), ps)                     # :267 serial, event order
    return img


def events_to_channels(xs, ys, ps, sensor_size=(180, 240)):
    """encodings.py:290-305.  [2,H,W] non-negative per-polarity counts.

    F9: the first events_to_image call zeroes out-of-range xs/ys of the CALLER
    but only a temporary ps, so the second call counts out-of-range negative
    events at row H-1, col 0.
    """
    xs, ys, ps = _f32(xs), _f32(ys), _f32(ps)
    assert len(xs) == len(ys) and len(ys) == len(ps)
    mask_pos = ps.copy()
    mask_neg = ps.copy()
    mask_pos[ps < 0] = 0
    mask_neg[ps > 0] = 0
    pos = events_to_image(xs, ys, ps * mask_pos, sensor_size)
    neg = events_to_image(xs, ys, ps * mask_neg, sensor_size)
    return np.stack([pos, neg])


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """encodings.py:272-287.  Temporal-bilinear voxel grid, y-flipped, [B,H,W]."""
    xs, ys, ts, ps = _f32(xs), _f32(ys), _f32(ts), _f32(ps)
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    voxel = []
    ts = ts * F32(num_bins - 1)                      # :280 (new array, caller's ts kept)
    for b in range(num_bins):
        w = np.maximum(F32(0), F32(1.0) - np.abs(ts - F32(b)))   # :283
        voxel.append(events_to_image(xs, ys, ps * w, sensor_size))
    return np.stack(voxel)


def interpolate_to_image(pxs, pys, dxs, dys, weights, img):
    """encodings.py:6-13.  Four serial scatter-adds (bilinear splat)."""
    one = F32(1.0)
    np.add.at(img, (pys, pxs), weights * (one - dxs) * (one - dys))
    np.add.at(img, (pys, pxs + 1), weights * dxs * (one - dys))
    np.add.at(img, (pys + 1, pxs), weights * (one - dxs) * dys)
    np.add.at(img, (pys + 1, pxs + 1), weights * dxs * dys)


def events_to_image_torch(xs, ys, ps, device=None, sensor_size=(180, 240),
                          clip_out_of_range=True, interpolation=None, padding=True):
    """encodings.py:16-72.  No y-flip; mutates xs, ys AND ps (views write through)."""
    xs, ys, ps = _f32(xs), _f32(ys), _f32(ps)
    H, W = sensor_size
    _oor_zero_inplace(xs, ys, ps, sensor_size)
    if interpolation == 'bilinear' and padding:
        img_size = (H + 1, W + 1)
    else:
        img_size = (H, W)
    mask = np.ones(xs.shape, dtype=np.float32)
    if clip_out_of_range:
        clipx = img_size[1] if (interpolation is None and padding is False) else img_size[1] - 1
        clipy = img_size[0] if (interpolation is None and padding is False) else img_size[0] - 1
        mask = np.where(xs >= clipx, F32(0), F32(1)) * np.where(ys >= clipy, F32(0), F32(1))
    img = np.zeros(img_size, dtype=np.float32)
    if interpolation == 'bilinear':
        pxs = np.floor(xs)
        pys = np.floor(ys)
        dxs = xs - pxs
        dys = ys - pys
        pxs = _trunc_long(pxs * mask)
        pys = _trunc_long(pys * mask)
        masked_ps = ps * mask
        interpolate_to_image(pxs, pys, dxs, dys, masked_ps, img)
    else:
        np.add.at(img, (_trunc_long(ys), _trunc_long(xs)), ps)
    return img


def binary_search_torch_tensor(t, l, r, x, side='left'):
    """encodings.py:75-97.  Returns ANY index whose value equals x (F10)."""
    if r is None:
        r = len(t) - 1
    while l <= r:
        if t[l] == x:
            return l
        if t[r] == x:
            return r
        mid = l + (r - l) // 2
        midval = t[mid]
        if midval == x:
            return mid
        elif midval < x:
            l = mid + 1
        else:
            r = mid - 1
    if side == 'left':
        return l
    return r


def _bin_slices(ts, B):
    """encodings.py:172-178 -- float32 bin boundaries and [beg, end) slices."""
    dt = F32(F32(ts[-1] - ts[0]) + F32(1e-6))
    delta_t = F32(dt / F32(B))
    out = []
    for bi in range(B):
        tstart = F32(ts[0] + F32(delta_t * F32(bi)))
        tend = F32(tstart + delta_t)
        beg = binary_search_torch_tensor(ts, 0, len(ts) - 1, tstart)
        end = binary_search_torch_tensor(ts, 0, len(ts) - 1, tend, side='right') + 1
        out.append((beg, end))
    return out, dt


def _early_out(ts, B, sensor_size):
    # encodings.py:122-123,166-167,217-218 (note: [B,H,W] even for the polarity stack)
    if np.sum(ts, dtype=np.float32) == 0 or len(ts) <= 3:
        return np.zeros((B, sensor_size[0], sensor_size[1]), dtype=np.float32)
    return None


def _slice(a, beg, end):
    # python slice semantics incl. negative end (end = r + 1 can be 0 when r == -1)
    return a[beg:end]


def events_to_stack_polarity(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240)):
    """encodings.py:151-199.  [2,B,H,W] counts, no flip; boundary double count (F10)."""
    xs, ys, ts, ps = _f32(xs), _f32(ys), _f32(ts), _f32(ps)
    eo = _early_out(ts, B, sensor_size)
    if eo is not None:
        return eo
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    slices, _ = _bin_slices(ts, B)
    pos, neg = [], []
    for beg, end in slices:
        p = _slice(ps, beg, end)
        mask_pos = p.copy()
        mask_neg = p.copy()
        mask_pos[p < 0] = 0
        mask_neg[p > 0] = 0
        vp = events_to_image_torch(_slice(xs, beg, end), _slice(ys, beg, end), p * mask_pos,
                                   sensor_size=sensor_size, clip_out_of_range=False)
        vn = events_to_image_torch(_slice(xs, beg, end), _slice(ys, beg, end), p * mask_neg,
                                   sensor_size=sensor_size, clip_out_of_range=False)
        pos.append(vp)
        neg.append(vn)
    return np.stack([np.stack(pos), np.stack(neg)])


def events_to_stack_no_polarity(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240)):
    """encodings.py:202-238.  [B,H,W] signed counts; ps slices are views -> mutated."""
    xs, ys, ts, ps = _f32(xs), _f32(ys), _f32(ts), _f32(ps)
    eo = _early_out(ts, B, sensor_size)
    if eo is not None:
        return eo
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    slices, _ = _bin_slices(ts, B)
    bins = []
    for beg, end in slices:
        bins.append(events_to_image_torch(_slice(xs, beg, end), _slice(ys, beg, end),
                                          _slice(ps, beg, end), sensor_size=sensor_size,
                                          clip_out_of_range=False))
    return np.stack(bins)


def events_to_voxel_torch(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240),
                          temporal_bilinear=True):
    """encodings.py:100-148.  No flip; bilinear-in-time or hard bins (F10 search)."""
    xs, ys, ts, ps = _f32(xs), _f32(ys), _f32(ts), _f32(ps)
    eo = _early_out(ts, B, sensor_size)
    if eo is not None:
        return eo
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    slices, dt = _bin_slices(ts, B)
    t_norm = (ts - ts[0]) / dt * F32(B - 1)          # :129 sub, true div, mul
    bins = []
    for bi in range(B):
        if temporal_bilinear:
            w = np.maximum(F32(0), F32(1.0) - np.abs(t_norm - F32(bi)))
            weights = ps * w
            vb = events_to_image_torch(xs, ys, weights, sensor_size=sensor_size,
                                       clip_out_of_range=False)
        else:
            beg, end = slices[bi]
            vb = events_to_image_torch(_slice(xs, beg, end), _slice(ys, beg, end),
                                       _slice(ps, beg, end), sensor_size=sensor_size,
                                       clip_out_of_range=False)
        bins.append(vb)
    return np.stack(bins)
