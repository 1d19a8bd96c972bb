"""Event encoders -- drop-in for the forward encoders of the reference `dataloader/encodings.py`
(:6-305): same names, signatures, return shapes and in-place side effects, computed by the
sm_100a histogram kernels of libbmc_b200 (csrc/encode.cu) on CUDA float32 tensors.

Differences a caller can observe:
  * inputs must live on a CUDA device (the reference runs these inside CPU dataloader workers;
    there is no CPU fallback here) and must be contiguous float32 (the reference's input contract,
    base_dataset.py:24-31);
  * count-valued encodings are bit-exact; float-weighted ones (events_to_voxel*, weighted
    events_to_image*) accumulate in a different order than the reference's serial loop and agree
    to ~1e-6 relative;
  * NaN coordinates are dropped (the reference indexes with garbage);
  * `event_restore` and `stack2cnt` return tensors on the caller's CUDA device; the reference moves its argument
    to the host first (`.cpu()`, encodings.py:594,658) and returns CPU tensors -- append `.cpu()` for that;
  * the reference's `ts.sum() == 0` early-out of the stack / voxel_torch encoders is decided from the two end
    stamps first and a full scan only when both are zero (exact for any input, one 4-byte read-back per call).
Reproduced on purpose (bit-exact parity): out-of-range events are zeroed in the
caller's xs / ys (/ ps), the leak of such events into pixel (0,0) on later passes, and the double
counting of bin-boundary events by the any-equal binary search.
"""
import ctypes as C

import torch

from .. import _lib
from .._lib import check, lib, stream_ptr

__all__ = ['interpolate_to_image', 'events_to_image_torch', 'binary_search_torch_tensor',
           'events_to_voxel_torch', 'events_to_stack_polarity', 'events_to_stack_no_polarity',
           'events_to_image', 'events_to_voxel', 'events_to_channels', 'events_to_channels_windows',
           'python_event_redistribute_PolarityStack', 'python_event_redistribute_NoPolarityStack', 'stack2cnt',
           'event_restore', 'deterministic']

_MUT = _lib.ENC_MUTATE

# Bit-reproducible float encodings.  Count-valued encodings (events_to_channels, the stacks) are integer histograms
# and always reproducible.  The float-weighted ones (events_to_voxel*, events_to_image* with non-unit weights) add
# fp32 values with atomics, whose order changes from run to run (~1e-7 relative jitter).  With DETERMINISTIC = True
# (or inside `with deterministic():`) they accumulate in 64-bit fixed point instead (BMC_ENC_DETERMINISTIC): the same
# events give the same bits on every run, at roughly a third of the throughput.  The reference's function signatures
# have no room for such a switch, hence the module-level one.
DETERMINISTIC = False


class deterministic:
    """Context manager: `with deterministic(): events_to_voxel(...)`."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global DETERMINISTIC
        self.prev, DETERMINISTIC = DETERMINISTIC, self.on
        return self

    def __exit__(self, *exc):
        global DETERMINISTIC
        DETERMINISTIC = self.prev


def _det():
    return (_lib.ENC_DETERMINISTIC if DETERMINISTIC else 0) | (_lib.ENC_SPLIT_BINS if SPLIT_BINS else 0)


# Grids beyond one SM's shared memory (events_to_channels from 2 x 58,112 pixels up, events_to_voxel* from 29,056 bins
# up to 180x320 pixels) switch to the split-bins kernel at 2^24 events (csrc/encode.cu, role_kernel: a few CTAs stream
# the same events, each holding a share of the bins); SPLIT_BINS = True takes it at any event count
# (BMC_ENC_SPLIT_BINS) -- same results, used by the parity tests.
SPLIT_BINS = False


def _chk(*ts):
    n = None
    for t in ts:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.BmcError('encoders need CUDA tensors (no CPU fallback), got %s' % (
                t.device if isinstance(t, torch.Tensor) else type(t)))
        if t.dtype != torch.float32 or t.dim() != 1 or not t.is_contiguous():
            raise _lib.BmcError('encoders need contiguous 1-D float32 event tensors (base_dataset.py:24-31), got '
                                '%s %s' % (t.dtype, tuple(t.shape)))
        n = len(t) if n is None else n
    return n


_WS = {}       # (device index, stream) -> scratch tensor, grown on demand


def _workspace(dev, nbytes):
    """Scratch for one encoder call.  The library clears what it uses at the start of every call (on the call's
    stream), so one buffer per (device, stream) is reused instead of a fresh `torch.empty` per call; calls on the
    same stream are ordered, calls on different streams get different buffers.  (Call with `dev` current.)"""
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


_WS_BYTES = {}     # output elements -> bmc_encode_workspace_bytes


def _run(fn, out, *args):
    """Call an encoder entry with a scratch workspace sized for `out`.  The per-call host path matters for the
    reference's call pattern (one call per 1024-2048 event window): no device-context switch when the tensors' device is
    the current one, workspace size and buffer looked up, not queried."""
    n = out.numel()
    nbytes = _WS_BYTES.get(n)
    if nbytes is None:
        nbytes = _WS_BYTES[n] = lib().bmc_encode_workspace_bytes(n)
    dev = out.device
    if dev.index == torch.cuda.current_device():
        ws = _workspace(dev, nbytes)
        check(fn(*args, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes))
    else:
        with torch.cuda.device(dev):
            ws = _workspace(dev, nbytes)
            check(fn(*args, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes))
    return out


def _p(t):
    return C.c_void_p(t.data_ptr())


def events_to_image(xs, ys, ps, sensor_size=(180, 240)):
    """Accumulate events into a y-flipped [H,W] image (reference encodings.py:241-269).
    Out-of-range events are zeroed in xs, ys AND ps, as the reference does."""
    n = _chk(xs, ys, ps)
    h, w = sensor_size
    out = torch.empty(h, w, dtype=torch.float32, device=xs.device)
    flags = _lib.ENC_FLIP_Y | _MUT | _det()
    return _run(lambda *a: lib().bmc_encode_image(_p(xs), _p(ys), _p(ps), n, h, w, *a, flags, stream_ptr()), out)


def events_to_channels(xs, ys, ps, sensor_size=(180, 240)):
    """Two-channel event counters [2,H,W] (reference encodings.py:290-305) -- the encoder on the
    live inference / training path (h5dataset.py:518-526)."""
    assert len(xs) == len(ys) and len(ys) == len(ps)
    n = _chk(xs, ys, ps)
    h, w = sensor_size
    out = torch.empty(2, h, w, dtype=torch.float32, device=xs.device)
    flags = _MUT | (_lib.ENC_SPLIT_BINS if SPLIT_BINS else 0)
    return _run(lambda *a: lib().bmc_encode_channels(_p(xs), _p(ys), _p(ps), n, h, w, *a, flags, stream_ptr()), out)


def events_to_channels_windows(xs, ys, ps, offsets, sensor_size=(180, 240)):
    """Batched events_to_channels: window i = events [offsets[i], offsets[i+1]) -> out[i] = [2,H,W].
    (Not in the reference; it is the dataloader's per-window loop, h5dataset.py:261-316, in one launch.)"""
    _chk(xs, ys, ps)
    if offsets.dtype != torch.int64 or not offsets.is_cuda:
        raise _lib.BmcError('offsets must be a CUDA int64 tensor')
    h, w = sensor_size
    nw = len(offsets) - 1
    out = torch.empty(nw, 2, h, w, dtype=torch.float32, device=xs.device)
    with torch.cuda.device(xs.device):
        check(lib().bmc_encode_channels_windows(_p(xs), _p(ys), _p(ps), _p(offsets), nw, h, w, _p(out), _MUT,
                                                stream_ptr()))
    return out


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """Temporal-bilinear voxel grid [B,H,W], y-flipped (reference encodings.py:272-287)."""
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    n = _chk(xs, ys, ts, ps)
    h, w = sensor_size
    out = torch.empty(num_bins, h, w, dtype=torch.float32, device=xs.device)
    flags = _lib.ENC_FLIP_Y | _MUT | _det()
    return _run(lambda *a: lib().bmc_encode_voxel(_p(xs), _p(ys), _p(ts), _p(ps), n, num_bins, h, w, *a, flags,
                                                  stream_ptr()), out)


def interpolate_to_image(pxs, pys, dxs, dys, weights, img):
    """Bilinear splat of weighted points into `img` (reference encodings.py:6-13).  Helper kept for
    API completeness; events_to_image_torch(interpolation='bilinear') does this inside its kernel."""
    img.index_put_((pys, pxs), weights * (1.0 - dxs) * (1.0 - dys), accumulate=True)
    img.index_put_((pys, pxs + 1), weights * dxs * (1.0 - dys), accumulate=True)
    img.index_put_((pys + 1, pxs), weights * (1.0 - dxs) * dys, accumulate=True)
    img.index_put_((pys + 1, pxs + 1), weights * dxs * dys, accumulate=True)


def events_to_image_torch(xs, ys, ps, device=None, sensor_size=(180, 240), clip_out_of_range=True,
                          interpolation=None, padding=True):
    """Event image without y-flip (reference encodings.py:16-72).  `interpolation='bilinear'`
    (with padding) splats into a (H+1)x(W+1) image.  Mutates xs, ys, ps like the reference."""
    n = _chk(xs, ys, ps)
    h, w = sensor_size
    flags = _MUT | _det()
    if interpolation == 'bilinear':
        if not padding:
            raise NotImplementedError('bilinear interpolation without padding (unused by the reference callers)')
        flags |= _lib.ENC_BILINEAR
        out = torch.empty(h + 1, w + 1, dtype=torch.float32, device=xs.device)
    else:
        out = torch.empty(h, w, dtype=torch.float32, device=xs.device)
    out = _run(lambda *a: lib().bmc_encode_image(_p(xs), _p(ys), _p(ps), n, h, w, *a, flags, stream_ptr()), out)
    return out if device is None else out.to(device)


def binary_search_torch_tensor(t, l, r, x, side='left'):
    """The reference's any-equal binary search (encodings.py:75-97), element reads on the host.
    The stack encoders below evaluate the same search on the device instead."""
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
    return l if side == 'left' else r


def _early_out(ts, B, sensor_size, device):
    # reference encodings.py:122-123: [B,H,W] zeros when `ts.sum() == 0 or len(ts) <= 3`; for the sorted,
    # normalised ts >= 0 of base_dataset.py:30 the first test is "every stamp is zero".  Read the LAST stamp (one
    # 4-byte copy; the first one is always 0 after normalisation) and scan the array only when that is zero too.
    if len(ts) > 3 and (ts[-1].item() != 0 or not bool((ts == 0).all())):
        return None
    return torch.zeros([B, sensor_size[0], sensor_size[1]], device=device)


def _stack_call(xs, ys, ps, ts_all, first, B, sensor_size, polarity):
    """The stack encoders on events [first, first + len(xs)) of a recording with timestamps ts_all (the whole
    recording when first == 0 and len(xs) == len(ts_all)).

    The reference returns [B,H,W] zeros when `ts.sum() == 0 or len(ts) <= 3` (encodings.py:122-123,166-167,
    217-218) -- for the sorted, normalised ts >= 0 of base_dataset.py:30 the first test is "every stamp is zero".
    The return SHAPE depends on it, so the host has to know; instead of reading ts before the launch (a stream
    synchronisation with the GPU idle behind it) the kernel is launched first with BMC_ENC_SKIP_ZERO_ENDS -- a
    recording whose two end stamps are zero then yields zeros and touches no event -- and the one-word verdict is
    read back afterwards.  Only when it is set is ts scanned."""
    h, w = sensor_size
    n_total = len(ts_all)
    if n_total <= 3:
        return torch.zeros([B, h, w], device=xs.device)
    out = torch.empty(*((2, B, h, w) if polarity else (B, h, w)), dtype=torch.float32, device=xs.device)
    nbytes = lib().bmc_encode_workspace_bytes(out.numel())
    with torch.cuda.device(out.device):
        ws = _workspace(out.device, nbytes)

    def launch(flags):
        with torch.cuda.device(out.device):
            check(lib().bmc_encode_stack_shard(_p(xs), _p(ys), _p(ps), len(xs), _p(ts_all), n_total, int(first), B, h, w,
                                               int(polarity), _p(out), _p(ws), nbytes, flags, stream_ptr()))
    launch(_MUT | _lib.ENC_SKIP_ZERO_ENDS)
    off = lib().bmc_encode_stack_flag_offset(out.numel())
    if ws[off:off + 4].view(torch.int32).item():          # both end stamps are zero
        if bool((ts_all == 0).all()):
            return torch.zeros([B, h, w], device=xs.device)
        launch(_MUT)                                      # unsorted / negative stamps: not the early-out after all
    return out


def _stack(xs, ys, ts, ps, B, device, sensor_size, polarity):
    if device is None:
        device = xs.device
    _chk(xs, ys, ts, ps)
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    return _stack_call(xs, ys, ps, ts, 0, B, sensor_size, polarity).to(device)


def events_to_stack_polarity(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240)):
    """[2,B,H,W] per-polarity counts per time bin (reference encodings.py:151-199)."""
    return _stack(xs, ys, ts, ps, B, device, sensor_size, True)


def events_to_stack_no_polarity(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240)):
    """[B,H,W] signed counts per time bin (reference encodings.py:202-238)."""
    return _stack(xs, ys, ts, ps, B, device, sensor_size, False)


def events_to_voxel_torch(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240), temporal_bilinear=True):
    """Voxel grid without y-flip (reference encodings.py:100-148): bilinear in time, or hard bins
    through the binary search when temporal_bilinear=False."""
    if not temporal_bilinear:
        return _stack(xs, ys, ts, ps, B, device, sensor_size, False)
    if device is None:
        device = xs.device
    n = _chk(xs, ys, ts, ps)
    eo = _early_out(ts, B, sensor_size, device)
    if eo is not None:
        return eo
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    h, w = sensor_size
    out = torch.empty(B, h, w, dtype=torch.float32, device=xs.device)
    flags = _lib.ENC_TNORM | _MUT | _det()
    out = _run(lambda *a: lib().bmc_encode_voxel(_p(xs), _p(ys), _p(ts), _p(ps), n, B, h, w, *a, flags,
                                                 stream_ptr()), out)
    return out.to(device)


# ---------------------------------------------------------------------------------------------- inverse encoders
def _redistribute(event_stack, mode, polarity):
    if not isinstance(event_stack, torch.Tensor) or not event_stack.is_cuda or event_stack.dtype != torch.float32:
        raise _lib.BmcError('event stacks must be CUDA float32 tensors (no CPU fallback)')
    want = 5 if polarity else 4
    if event_stack.dim() != want:
        raise _lib.BmcError('expected a %d-D stack, got %s' % (want, tuple(event_stack.shape)))
    if mode not in ('linear', 'random'):
        raise ValueError(mode)
    st = event_stack.contiguous()
    b = st.shape[0]
    p = 2 if polarity else 1
    c, y, x = st.shape[-3:]
    if polarity and st.shape[1] != 2:
        raise _lib.BmcError('polarity stacks are [B, 2, C, Y, X]')
    per_entry = p * c * y * x
    dev = st.device
    with torch.cuda.device(dev):
        nb0 = lib().bmc_stack_to_events_workspace_bytes(b, per_entry, 0)
        ws0 = torch.empty(nb0, dtype=torch.uint8, device=dev)
        counts = torch.empty(2, b, dtype=torch.int64, device=dev)
        check(lib().bmc_stack_event_counts(_p(st), b, per_entry, _p(ws0), nb0, _p(counts[0]), _p(counts[1]), stream_ptr()))
        totals, sums = counts.tolist()                      # the output SHAPE depends on them (reference: maxlen, :405-408)
        if sum(sums) == 0:                                  # `if event_stack.sum() != 0` (:380,429)
            return torch.zeros(b, 1, 4, device=dev)
        maxlen = max(t if s != 0 else 1 for t, s in zip(totals, sums))
        nb = lib().bmc_stack_to_events_workspace_bytes(b, per_entry, maxlen)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        out = torch.empty(b, maxlen, 4, dtype=torch.float32, device=dev)
        rnd = torch.rand(b, maxlen, device=dev) if mode == 'random' else None
        check(lib().bmc_stack_to_events(_p(st), b, p, c, y, x, maxlen, _p(rnd) if rnd is not None else None, _p(out),
                                        _p(ws), nb, stream_ptr()))
    return out


def python_event_redistribute_PolarityStack(event_stack, mode='linear'):
    """[B, 2, C, Y, X] count stack -> batched event cloud [B, max_num_event, 4] = (x, y, t, p), each entry stably
    sorted by t and zero-padded (reference encodings.py:367-414).  mode='random' draws its own uniform numbers
    (torch.rand on the device), so only its distribution matches the reference."""
    return _redistribute(event_stack, mode, True)


def python_event_redistribute_NoPolarityStack(event_stack, mode='linear'):
    """[B, C, Y, X] signed count stack -> batched event cloud [B, max_num_event, 4] (reference encodings.py:417-464)."""
    return _redistribute(event_stack, mode, False)


def stack2cnt(stack):
    """[B, TB, H, W] signed stack -> [B, 2, H, W] positive / negative counts (reference encodings.py:653-671)."""
    if not isinstance(stack, torch.Tensor) or not stack.is_cuda or stack.dtype != torch.float32 or stack.dim() != 4:
        raise _lib.BmcError('stack2cnt needs a 4-D CUDA float32 tensor (no CPU fallback)')
    st = stack.contiguous()
    b, tb, h, w = st.shape
    out = torch.empty(b, 2, h, w, dtype=torch.float32, device=st.device)
    with torch.cuda.device(st.device):
        check(lib().bmc_stack2cnt(_p(st), b, tb, h, w, _p(out), stream_ptr()))
    return out


def event_restore(events, resolution):
    """[B, N, 4] normalised (x, y, t, p) -> pixel coordinates and +-1 polarities (reference encodings.py:581-602);
    elementwise tensor glue, stays on the caller's device."""
    events = events.detach().clone()
    x = events[:, :, 0] * resolution[1]
    y = events[:, :, 1] * resolution[0]
    t = events[:, :, 2]
    p = torch.sign(events[:, :, 3])
    return torch.stack([x, y, t, p], dim=2)
