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
