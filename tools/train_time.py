"""Time the BMCNet training iteration (BASELINE config 5) eager vs GraphedIteration at a few batch sizes.
usage: python tools/train_time.py [batches ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [2, 8]
    cx = bench.Ctx()
    cx.dev = torch.device('cuda', 0)
    cx.world, cx.rank, cx.models, cx.sampler = 1, 0, {}, None
    torch.cuda.set_device(0)
    for b in batches:
        r = bench.run_train(cx, batch=b, iters=3)
        print(json.dumps({k: r[k] for k in ('batch_per_gpu', 'ms_per_iteration', 'ms_per_iteration_eager', 'value', 'loss_finite')}))
        print('peak mem GB', torch.cuda.max_memory_allocated() / 1e9)
        torch.cuda.reset_peak_memory_stats()


if __name__ == '__main__':
    main()
