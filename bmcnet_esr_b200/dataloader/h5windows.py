"""The reference dataloader's window pipeline (dataloader/h5dataset.py, dataloader/base_dataset.py) on the
device, for recordings that are already resident in HBM as stored on disk (int16 xs / ys, float64 ts / ps,
generate_dataset/tools/event_packagers.py:128-156).

The reference reads each window from HDF5 in a DataLoader worker, casts it (`event_formatting`), encodes it on
the CPU (`create_cnt_encoding`) and ships the count frames to the GPU; `SequenceDataset` re-encodes every window
up to `seql` times.  Here every window of a recording is encoded exactly once, in one launch."""
import ctypes as C

import torch

from .. import _lib

__all__ = ['compute_k_indices', 'event_formatting', 'windows_to_counts', 'sequence_tuples', 'sequence_item']


def compute_k_indices(num_events, window, sliding_window, dataset_length=None):
    """[[idx0, idx1], ...] of H5Dataset.compute_k_indices (h5dataset.py:197-210) with the length rule of
    H5Dataset.__init__ (`data_mode == 'events'`, h5dataset.py:169-175)."""
    stride = window - sliding_window
    length = max(int(num_events / stride), 0)
    if dataset_length is not None and dataset_length <= length:
        length = dataset_length
    out = []
    for i in range(length):
        idx0 = stride * i
        idx1 = min(idx0 + window, num_events - 1)
        out.append([idx0, idx1])
    return out


def _raw(xs, ys, ts, ps):
    for t, dt in ((xs, torch.int16), (ys, torch.int16), (ts, torch.float64), (ps, torch.float64)):
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != dt or t.dim() != 1 or not t.is_contiguous():
            raise _lib.BmcError('raw recordings are contiguous 1-D CUDA tensors: int16 xs / ys, float64 ts / ps '
                                '(event_packagers.py:128-156); no CPU fallback')


def event_formatting(events):
    """BaseDataset.event_formatting (base_dataset.py:24-31): (xs, ys, ts, ps) raw CUDA tensors -> float32 [4, N]
    with ts normalised to [0, 1)."""
    xs, ys, ts, ps = events
    _raw(xs, ys, ts, ps)
    n = len(xs)
    assert len(ys) == n and len(ts) == n and len(ps) == n
    out = torch.empty(4, n, dtype=torch.float32, device=xs.device)
    with torch.cuda.device(xs.device):
        _lib.check(_lib.lib().bmc_format_events(C.c_void_p(xs.data_ptr()), C.c_void_p(ys.data_ptr()), C.c_void_p(ts.data_ptr()),
                                                C.c_void_p(ps.data_ptr()), n, C.c_void_p(out.data_ptr()), _lib.stream_ptr()))
    return out


def windows_to_counts(xs, ys, ps, window, sliding_window, sensor_size, dataset_length=None):
    """Count frames [n_windows, 2, H, W] of every window of a recording: `H5Dataset.__getitem__(i)['inp_cnt']`
    (h5dataset.py:261-316) for i in range(len(dataset)), in one launch."""
    _raw(xs, ys, None, ps)
    n = len(xs)
    assert len(ys) == n and len(ps) == n
    stride = window - sliding_window
    n_win = len(compute_k_indices(n, window, sliding_window, dataset_length))
    h, w = sensor_size
    out = torch.empty(n_win, 2, h, w, dtype=torch.float32, device=xs.device)
    if n_win == 0:                       # recording shorter than one stride (the reference raises on length 0, h5dataset.py:194-195)
        return out
    with torch.cuda.device(xs.device):
        _lib.check(_lib.lib().bmc_encode_channels_windows_raw(C.c_void_p(xs.data_ptr()), C.c_void_p(ys.data_ptr()),
                                                              C.c_void_p(ps.data_ptr()), n, window, stride, n_win, h, w,
                                                              C.c_void_p(out.data_ptr()), 0, _lib.stream_ptr()))
    return out


def sequence_tuples(cnt, seqn=2):
    """Sliding `seqn`-tuples of consecutive count frames as a VIEW: [n_win, 2, H, W] -> [n_win - seqn + 1, seqn, 2, H, W]
    (what `custom_collate` + `concat_dict` stack for batch 1, h5dataloader.py:293-330; no copy, no re-encoding)."""
    n = cnt.shape[0] - seqn + 1
    if n <= 0:
        raise ValueError('need at least %d windows, got %d' % (seqn, cnt.shape[0]))
    st = cnt.stride()
    return cnt.as_strided((n, seqn) + tuple(cnt.shape[1:]), (st[0], st[0]) + tuple(st[1:]))


def sequence_item(cnt, i, sequence_length, seqn=2, step_size=1):
    """Item i of the reference's InferenceHDF5DataLoaderSequence for one recording (batch 1): the list of
    `sequence_length - seqn + 1` dicts-worth of 'inp_cnt' tensors [1, seqn, 2, H, W] built from windows
    [i*step_size, i*step_size + sequence_length) (SequenceDataset.__getitem__, h5dataset.py:666-700; custom_collate,
    h5dataloader.py:293-317).  The inference loop reads element 0 (infer_BMCNet.py:48)."""
    j = i * step_size
    if j + sequence_length > cnt.shape[0]:
        raise IndexError('item %d needs windows up to %d, the recording has %d' % (i, j + sequence_length, cnt.shape[0]))
    t = sequence_tuples(cnt[j:j + sequence_length], seqn)
    return [t[m:m + 1] for m in range(t.shape[0])]
