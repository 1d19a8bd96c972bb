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
