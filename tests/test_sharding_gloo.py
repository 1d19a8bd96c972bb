"""World-size-2 gloo test of the multi-GPU host logic (replicas over independent recordings)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bmcnet_esr_b200.sharding import frames_total, shard_sequences


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_seq, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_sequences(n_seq, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    total = frames_total(len(mine) * 3)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((gathered, total, t.item()))
    dist.destroy_process_group()


def test_round_robin_sharding_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, total, tmax = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert gathered == [[0, 2, 4, 6], [1, 3, 5]]
    assert sorted(sum(gathered, [])) == list(range(7))
    assert total == 21 and tmax == 2.0


def test_shard_edges():
    assert shard_sequences(0, 0, 4) == []
    assert shard_sequences(3, 3, 4) == []
    assert sum((shard_sequences(10, r, 4) for r in range(4)), []).__len__() == 10


# ---- one recording split into contiguous event ranges: partial grids + one all-reduce ----

def _events(n, h, w, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    xs = rng.integers(-2, w + 2, n).astype(np.float32)        # a few out-of-range events (quirk F9)
    ys = rng.integers(-1, h + 1, n).astype(np.float32)
    ts = np.sort(rng.random(n)).astype(np.float32)
    ps = (rng.integers(0, 2, n) * 2 - 1).astype(np.float32)
    return xs, ys, ts, ps


def _grid_worker(rank, world, port, n, out):
    import numpy as np
    from oracle import encodings_np as O
    from bmcnet_esr_b200.sharding import event_range, sum_grids
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    h, w = 12, 20
    xs, ys, ts, ps = _events(n, h, w, 5)
    lo, hi = event_range(n, rank, world)
    cnt = torch.from_numpy(O.events_to_channels(xs[lo:hi].copy(), ys[lo:hi].copy(), ps[lo:hi].copy(), (h, w)))
    vox = torch.from_numpy(O.events_to_voxel(xs[lo:hi].copy(), ys[lo:hi].copy(), ts[lo:hi].copy(), ps[lo:hi].copy(), 5, (h, w)))
    sum_grids(cnt)
    sum_grids(vox)
    if rank == 0:
        out.put((cnt.numpy(), vox.numpy()))
    dist.destroy_process_group()


def test_split_recording_partial_grids_world2():
    import numpy as np
    from oracle import encodings_np as O
    n = 4099
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grid_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    cnt, vox = q.get(timeout=180)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    xs, ys, ts, ps = _events(n, 12, 20, 5)
    ref_cnt = O.events_to_channels(xs.copy(), ys.copy(), ps.copy(), (12, 20))
    ref_vox = O.events_to_voxel(xs.copy(), ys.copy(), ts.copy(), ps.copy(), 5, (12, 20))
    assert np.array_equal(cnt, ref_cnt)                       # integer counts: exact in any order
    assert np.abs(vox - ref_vox).max() <= 1e-6 * max(1.0, np.abs(ref_vox).max())


def test_event_range_covers_once_and_stays_aligned():
    from bmcnet_esr_b200.sharding import event_range
    for n in (0, 1, 3, 17, 4096, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [event_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all(lo <= hi and (lo % 4 == 0) for lo, hi in cuts)


# ---- data-parallel training step: one all-reduce of the alias-deduplicated gradients ----
def _grad_worker(rank, world, port, out):
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models._train import allreduce_gradients, unique_parameters
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    m = BMCNet(4, 128, 5)
    params = unique_parameters(m)
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = allreduce_gradients(params)
    if rank == 0:
        out.put((n, len(params), [p.grad.flatten()[0].item() for p in params[:3]], len(m.state_dict())))
    dist.destroy_process_group()


def test_gradient_allreduce_is_alias_deduplicated_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n, n_params, first, n_keys = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert n == 2731680 and n_params == 54 and n_keys == 318       # 318 keys, 54 Parameters, 2,731,680 elements
    assert first == [1.5, 3.0, 4.5]                                # mean over the two ranks of (rank+1)*(i+1)
