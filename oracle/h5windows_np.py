"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product) of the reference dataloader's window
pipeline: H5Dataset.compute_k_indices (dataloader/h5dataset.py:169-175, 197-210), get_events (:407-414),
BaseDataset.event_formatting (dataloader/base_dataset.py:24-31) and create_cnt_encoding (:518-526).

`event_formatting` is pinned against the reference function itself (tests/golden/fmt_events.npz, written by
oracle/make_golden.py); h5dataset.py cannot be imported here (h5py is absent), so the window indexing is a
restatement checked by its defining properties in tests/test_oracle_vs_golden.py."""
import numpy as np

from . import encodings_np as E


def compute_k_indices(num_events, window, sliding_window, dataset_length=None):
    stride = window - sliding_window
    max_length = max(int(num_events / stride), 0)                       # h5dataset.py:170
    length = max_length
    if dataset_length is not None:                                      # :171-172
        length = dataset_length if dataset_length <= max_length else max_length
    k_indices = []
    for i in range(length):                                             # :205-210
        idx0 = stride * i
        idx1 = idx0 + window
        if idx1 > num_events - 1:
            idx1 = num_events - 1
        k_indices.append([idx0, idx1])
    return k_indices


def event_formatting(events):
    """events: (xs, ys, ts, ps) numpy arrays of any dtype -> float32 [4, N] (base_dataset.py:24-31)."""
    xs, ys, ts, ps = (np.asarray(e).astype(np.float32) for e in events)
    ts = (ts - ts[0]) / np.float32(np.float32(ts[-1] - ts[0]) + np.float32(1e-6))
    return np.stack([xs, ys, ts.astype(np.float32), ps])


def windows_to_counts(xs, ys, ts, ps, window, sliding_window, sensor_size, dataset_length=None):
    """inp_cnt of every dataset item (h5dataset.py:261-316), window by window like the reference."""
    out = []
    for idx0, idx1 in compute_k_indices(len(xs), window, sliding_window, dataset_length):
        ev = np.concatenate((xs[np.newaxis, idx0:idx1], ys[np.newaxis, idx0:idx1], ts[np.newaxis, idx0:idx1],
                             ps[np.newaxis, idx0:idx1]), axis=0)        # get_events, :407-414 (promotes to float64)
        f = event_formatting(ev)
        out.append(E.events_to_channels(f[0].copy(), f[1].copy(), f[3].copy(), sensor_size=sensor_size))
    return np.stack(out) if out else np.zeros((0, 2) + tuple(sensor_size), np.float32)
