"""One recording encoded by two ranks (SURVEY.md 8e "one giant grid"): each rank runs the CUDA encoder on its
contiguous event range and one all-reduce adds the partial grids (bmcnet_esr_b200/sharding.py).

Two processes: NCCL with one GPU each when the box has two GPUs, otherwise both ranks share cuda:0 and the
all-reduce goes through gloo on the CUDA tensors (NCCL refuses two ranks on one device) -- the encoder kernels
and the range / boundary logic are the same either way.  Bars: counts and stacks bit-exact against the CPU
oracle on the whole recording, voxels within 1e-6 of the largest voxel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
H, W, B, N = 45, 80, 5, 200_003


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _recording(n=N):
    rng = np.random.default_rng(11)
    xs = rng.integers(-1, W + 1, n).astype(np.float32)          # some out-of-range events (quirk F9)
    ys = rng.integers(0, H, n).astype(np.float32)
    ts = np.sort(rng.random(n)).astype(np.float32)
    ts[n // 3:n // 3 + 6] = ts[n // 3]                          # duplicate timestamps
    ts = ((ts - ts[0]) / (ts[-1] - ts[0] + np.float32(1e-6))).astype(np.float32)
    ps = (rng.integers(0, 2, n) * 2 - 1).astype(np.float32)
    return xs, ys, ts, ps


def _worker(rank, world, port, n_gpus, out):
    from bmcnet_esr_b200 import sharding as S
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank if n_gpus >= world else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl' if n_gpus >= world else 'gloo', rank=rank, world_size=world)
    xs, ys, ts, ps = _recording()
    lo, hi = S.event_range(len(xs), rank, world)
    cut = lambda a: torch.from_numpy(a[lo:hi].copy()).to(dev)
    ts_all = torch.from_numpy(ts).to(dev)
    res = {
        'channels': S.events_to_channels_sharded(cut(xs), cut(ys), cut(ps), sensor_size=(H, W)),
        # in-range events for the float-weighted grid: thousands of out-of-range events piling onto one pixel
        # would make the ORACLE's serial fp32 sum the inexact side (the quirk itself is covered by the counts)
        'voxel': S.events_to_voxel_sharded(cut(np.clip(xs, 0, W - 1)), cut(ys), cut(ts), cut(ps), B, sensor_size=(H, W)),
        'stack_polarity': S.events_to_stack_polarity_sharded(cut(xs), cut(ys), ts_all, cut(ps), lo, B, sensor_size=(H, W)),
        'stack_no_polarity': S.events_to_stack_no_polarity_sharded(cut(xs), cut(ys), ts_all, cut(ps), lo, B,
                                                                   sensor_size=(H, W)),
        # the early-out is a property of the whole recording: [B,H,W] zeros on every rank
        'early_out': S.events_to_stack_polarity_sharded(cut(xs), cut(ys), torch.zeros_like(ts_all), cut(ps), lo, B,
                                                        sensor_size=(H, W)),
    }
    torch.cuda.synchronize()
    out.put((rank, {k: v.cpu().numpy() for k, v in res.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_encode_one_recording():
    from oracle import encodings_np as O
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, torch.cuda.device_count(), q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    xs, ys, ts, ps = _recording()
    ref = {
        'channels': O.events_to_channels(xs.copy(), ys.copy(), ps.copy(), (H, W)),
        'stack_polarity': O.events_to_stack_polarity(xs.copy(), ys.copy(), ts.copy(), ps.copy(), B, sensor_size=(H, W)),
        'stack_no_polarity': O.events_to_stack_no_polarity(xs.copy(), ys.copy(), ts.copy(), ps.copy(), B,
                                                           sensor_size=(H, W)),
    }
    vox = O.events_to_voxel(np.clip(xs, 0, W - 1), ys.copy(), ts.copy(), ps.copy(), B, (H, W))
    for rank in range(world):                       # the all-reduce leaves the full grid on every rank
        for k, r in ref.items():
            assert got[rank][k].shape == r.shape, k
            assert np.array_equal(got[rank][k], r), 'rank %d %s differs from the oracle' % (rank, k)
        assert np.abs(got[rank]['voxel'] - vox).max() <= 1e-6 * np.abs(vox).max()
        assert got[rank]['early_out'].shape == (B, H, W) and not got[rank]['early_out'].any()


def test_shard_entry_matches_whole_recording_on_one_gpu():
    """bmc_encode_stack_shard over two ranges, summed by hand, equals bmc_encode_stack on all events -- incl. a
    range that is empty and a cut that is not 16-byte aligned."""
    from bmcnet_esr_b200 import sharding as S
    from bmcnet_esr_b200.dataloader import encodings as G
    xs, ys, ts, ps = _recording(50_001)
    dev = torch.device('cuda')
    full = [torch.from_numpy(a.copy()).to(dev) for a in (xs, ys, ts, ps)]
    want = G.events_to_stack_polarity(full[0], full[1], full[2], full[3], B, sensor_size=(H, W))
    ts_all = torch.from_numpy(ts).to(dev)
    for cuts in ((0, 20_001, 50_001), (0, 0, 50_001), (0, 50_001, 50_001)):
        acc = torch.zeros_like(want)
        for lo, hi in zip(cuts, cuts[1:]):
            cut = lambda a: torch.from_numpy(a[lo:hi].copy()).to(dev)
            acc += S.events_to_stack_polarity_sharded(cut(xs), cut(ys), ts_all, cut(ps), lo, B, sensor_size=(H, W))
        assert torch.equal(acc, want), cuts
    with pytest.raises(Exception):
        S.events_to_stack_polarity_sharded(full[0], full[1], ts_all[:100], full[3], 0, B, sensor_size=(H, W))
