"""Which recordings a rank runs.  Inference is a sequential recurrence per recording and recordings are
independent (reference infer_BMCNet.py:260-282 loops over `data_list`), so multi-GPU inference is
"replicas only": rank r of `world` takes recordings r, r + world, ... and no data-path collective exists."""


def shard_sequences(n_sequences, rank, world):
    """Indices of the recordings rank `rank` processes (round-robin, covers every index exactly once)."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return list(range(rank, n_sequences, world))


def frames_total(frames_per_rank, group=None):
    """Whole-job frame count: sum of the per-rank counts (the one reduction the bench needs)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(frames_per_rank)], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


# ---------------------------------------------------------------------------------------------------------
# One recording encoded by several GPUs ("one giant grid").  The encodings are sums over
# events, so rank r encodes the contiguous range event_range(n, r, world) into a partial grid and ONE
# all-reduce (NCCL over NVLink on the GPUs; 28.8 KB .. 9.2 MB against gigabytes of events) adds them.
# Counts and stacks are integer-valued floats below 2^24, so the sum is exact in any order and the result is
# bit-identical to the single-GPU encoder; voxel weights are fp32 partial sums (same 1e-6 relative bound).
# ---------------------------------------------------------------------------------------------------------

def event_range(n_events, rank, world, align=4):
    """[lo, hi) of the events rank `rank` encodes: contiguous, near-equal, covering [0, n) exactly once; interior
    cuts fall on multiples of `align` events so every rank's slice keeps the 16-byte alignment of the
    vectorised loads."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))

    def cut(r):
        if r >= world:
            return n_events
        return min(n_events, (n_events * r // world) // align * align)
    return cut(rank), cut(rank + 1)


def sum_grids(grid, group=None):
    """In-place sum of the ranks' partial grids (the path's only exchange step); identity without a process
    group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
    return grid


def events_to_channels_sharded(xs, ys, ps, sensor_size=(180, 240), group=None):
    """events_to_channels (reference encodings.py:290-305) of a recording split across ranks: pass this rank's
    slice; every rank gets the full [2,H,W] counts."""
    from .dataloader.encodings import events_to_channels
    return sum_grids(events_to_channels(xs, ys, ps, sensor_size=sensor_size), group)


def events_to_voxel_sharded(xs, ys, ts, ps, num_bins, sensor_size=(180, 240), group=None):
    """events_to_voxel (reference encodings.py:272-287) of a split recording; ts is already normalised over the
    WHOLE recording (base_dataset.py:30), so an event's bin weights do not depend on the split."""
    from .dataloader.encodings import events_to_voxel
    return sum_grids(events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=sensor_size), group)


def _stack_sharded(xs, ys, ts_all, ps, first, B, sensor_size, polarity, group):
    from . import _lib
    from .dataloader import encodings as E
    n_local = E._chk(xs, ys, ps)
    E._chk(ts_all)
    if first < 0 or first + n_local > len(ts_all):
        raise _lib.BmcError('events [%d, %d) outside the recording of %d' % (first, first + n_local, len(ts_all)))
    out = E._stack_call(xs, ys, ps, ts_all, first, B, sensor_size, polarity)
    if polarity and out.dim() == 3:        # the early-out: a property of the whole recording, same on every rank
        return out
    return sum_grids(out, group)           # (the no-polarity early-out is [B,H,W] zeros on every rank: summing is harmless)


def events_to_stack_polarity_sharded(xs, ys, ts_all, ps, first, B, sensor_size=(180, 240), group=None):
    """events_to_stack_polarity (reference encodings.py:151-199) of a split recording.  xs, ys, ps: this rank's
    events [first, first + len(xs)); ts_all: the timestamps of the WHOLE recording (the bin boundaries and the
    any-equal binary search of encodings.py:75-97 are evaluated on it by every rank)."""
    return _stack_sharded(xs, ys, ts_all, ps, first, B, sensor_size, True, group)


def events_to_stack_no_polarity_sharded(xs, ys, ts_all, ps, first, B, sensor_size=(180, 240), group=None):
    """events_to_stack_no_polarity (reference encodings.py:202-238) of a split recording; see the polarity form."""
    return _stack_sharded(xs, ys, ts_all, ps, first, B, sensor_size, False, group)
