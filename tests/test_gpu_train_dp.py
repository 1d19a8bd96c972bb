"""Data-parallel training step on two ranks (BASELINE config 5; SURVEY 8e "Training"): every rank runs the
sequence forward + BPTT on ITS sequences through the kernels, ONE all-reduce averages the alias-deduplicated
gradients (models/_train.py allreduce_gradients over the flat buffer of FusedAdamAMSGrad), every rank applies the
same fused Adam step.  NCCL with one GPU per rank when the box has two GPUs; otherwise both ranks share cuda:0 and
the all-reduce goes through gloo (NCCL refuses two ranks on one device) -- kernels and host logic are identical.

Bar: the averaged gradients equal the fp32 oracle's gradients of the COMBINED batch (mean-reduced MSE: the mean of
the per-rank gradients) within 1.5 % of each parameter's max-abs gradient, and the two ranks end the step with
bit-identical parameters."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
H, W, STEPS = 16, 24, 3


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    from oracle.make_golden import synth_counts
    xs = [synth_counts(2, H, W, 900 + s) for s in range(STEPS)]
    g = torch.Generator().manual_seed(9)
    gts = [torch.poisson(torch.full((2, 2, 4 * H, 4 * W), 0.3), generator=g) for _ in range(STEPS)]
    return xs, gts


def _worker(rank, world, port, n_gpus, ckpt, out):
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, allreduce_gradients
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank if n_gpus >= world else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl' if n_gpus >= world else 'gloo', rank=rank, world_size=world)
    sd = torch.load(ckpt, map_location='cpu')
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    opt = FusedAdamAMSGrad(m.parameters())
    xs, gts = _data()
    opt.zero_grad()
    st = [torch.zeros(1, 128, H, W, device=dev), torch.zeros(1, 32, H, W, device=dev)]
    loss, init = 0, True
    for x, gt in zip(xs, gts):                                    # this rank's sequence: sample `rank` of the batch
        st = list(m(x[rank:rank + 1].to(dev), *st, init))
        init = False
        loss = loss + F.mse_loss(st[-1], gt[rank:rank + 1].to(dev))
    loss.backward()
    n = allreduce_gradients(opt)
    grads = {k: p.grad.detach().cpu().numpy().copy() for k, p in m.named_parameters()}
    opt.step()
    torch.cuda.synchronize()
    # numpy, not tensors: a tensor travels through the queue as a shared-memory handle that dies with this process
    out.put((rank, n, loss.item(), grads, {k: p.detach().cpu().numpy().copy() for k, p in m.named_parameters()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step(plain_ckpt):
    from oracle import bmcnet_fp32 as O
    from oracle import train_step as T
    CKPT_PLAIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref', 'BMCNet_plain_nfs_x4.pth')
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, torch.cuda.device_count(), CKPT_PLAIN, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, n, loss, grads, params = q.get(timeout=600)
        got[r] = (n, loss, grads, params)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got[0][0] == got[1][0] == 1003296                     # unique (alias-deduplicated) elements, SURVEY F4
    xs, gts = _data()
    ref_loss, ref_grads, _ = T.loss_and_grads(plain_ckpt, xs, gts, True)
    assert abs(0.5 * (got[0][1] + got[1][1]) - ref_loss.item()) <= 1e-3 * ref_loss.item()
    for k, g in got[0][2].items():
        r = ref_grads[O._alias_root(k)].numpy()
        assert np.array_equal(g, got[1][2][k]), k                # both ranks hold the same reduced gradient
        assert np.abs(g - r).max() <= 1.5e-2 * np.abs(r).max(), k
    for k, p in got[0][3].items():
        assert np.array_equal(p, got[1][3][k]), k                # replicas stay bit-identical after the step
