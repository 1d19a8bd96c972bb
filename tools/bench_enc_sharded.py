"""Aggregate rate of ONE recording encoded by all ranks (SURVEY.md 8e "one giant grid"): contiguous event range
per rank -> partial grid -> one NCCL all-reduce.  Launch: python -m torch.distributed.run --nnodes=1
--nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_enc_sharded.py [--events-per-gpu 4e8].
Device-timed (CUDA events), max over ranks; rank 0 prints one JSON line per encoder."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('NCCL_DEBUG', 'WARN')
from bmcnet_esr_b200 import sharding as S      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--events-per-gpu', type=float, default=4e8)
ap.add_argument('--reps', type=int, default=5)
args = ap.parse_args()
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
h, w, B = 45, 80, 5
n = int(args.events_per_gpu)
g = torch.Generator(device=dev).manual_seed(100 + rank)
xs = torch.rand(n, device=dev, generator=g) * w
ys = torch.rand(n, device=dev, generator=g) * h
ps = (torch.rand(n, device=dev, generator=g) < 0.5).float() * 2 - 1
ts = (torch.sort(torch.rand(n, device=dev, generator=g))[0] + rank) / world      # rank r holds [r/world, (r+1)/world)
runs = {
    'events_to_channels': (12.0, lambda: S.events_to_channels_sharded(xs, ys, ps, sensor_size=(h, w))),
    'events_to_voxel': (16.0, lambda: S.events_to_voxel_sharded(xs, ys, ts, ps, B, sensor_size=(h, w))),
}
for name, (bpe, fn) in runs.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = float(out.double().abs().sum())
    if rank == 0:
        sec = ms.item() * 1e-3
        print(json.dumps({'encoder': name, 'n_gpus': world, 'events_total': n * world, 'ms': ms.item(),
                          'gevents_per_s': n * world / sec / 1e9, 'algorithmic_GBps': bpe * n * world / sec / 1e9,
                          'grid_abs_sum': tot, 'scaling': 'weak'}))
if world > 1:
    dist.destroy_process_group()
