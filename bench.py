"""Headline benchmark: x4 SR event-frames/s of the BMCNet hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload plain_nfs|bmcnet_nfs|bmcnet_eventzoom]
                    [--batch B] [--impl reference] [--only-headline]

One STEP = one recurrent step of the hot path for a batch of B independent synthetic sequences on
each GPU: encode the step's two event windows per sequence into per-polarity count frames
(`events_to_channels`, the encoder on the live path) and run one `forward` of the model
(x4 prediction for every sequence).  value = event-frames/s summed over all GPUs.

The headline line is the workload given by --workload (default plain_nfs = BASELINE configs[1], the one
checkpoint the reference ships).  Unless --only-headline is given, the SAME line also carries, under
"workloads", the identically measured numbers (value / e2e / ms_per_step / clocks / sustained) of the other
two: `bmcnet_nfs` (configs[0]'s model and shape at a batch that fills the GPU) and `bmcnet_eventzoom`
(configs[3]: batched sequences sharded over 1/2/4/8 GPUs) -- so a driver that only ever runs the default
command still records the model the north star is about, at every N.

  value   device-resident: events already in HBM, recurrent state kept in the arena
          (bmc_model_step), prediction written to HBM.  EXACTLY K timed steps.
  value_sustained  the same loop run for >= 2 s right after (K-step regions of ~80 ms are a clock burst).
  e2e     through the reference-facing API (`model(x, h, o, init)` + `events_to_channels_windows`)
          with the step's events copied from pinned host memory and the prediction read back to
          the host inside the timed region, every step; timed 3 x K steps, the median run is reported
          (all three are listed: the host side of a shared box is noisy).
  latency_b1  batch-1 latency, the shape infer_BMCNet.py:46-68 runs: CUDA events around `forward` alone
          (exactly its starter/ender pair), median ms per frame, for BMCNet_plain and BMCNet at 45x80.
  eager_b200  the UNMODIFIED reference modules (baseline/_ref) moved to the same B200 in eager mode (fp32, stock
          cuDNN/cuBLAS settings): the "existing Blackwell path" -- batch 1 latency and batched frames/s.
  roofline  the dominant kernel (3x3 128->128 implicit-GEMM conv, tcgen05: conv_slab2_tc) timed live,
          back to back, at this workload's shape; algorithmic FLOPs = 2*147456 MAC per real LR pixel;
          `traffic` = DRAM bytes per launch of the committed ncu --set full capture (profiles/).
  cpu_baseline / --impl reference  the reference's own CPU path on all host cores: its modules and its
          `events_to_channels` from baseline/_ref (kind "reference"; the oracle port, kind "port", only when the
          tree was not staged), on a bounded sample (8 sequences per step).

Inference shards by independent sequences: every rank runs its own batch, no collective on the
data path ("scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints ONE JSON line on stdout; NCCL's version banner (NCCL_DEBUG=VERSION) would precede it
if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
    os.environ['NCCL_DEBUG'] = 'WARN'

WORKLOADS = {
    # name: (model, LR H, W, events per window, description)
    'plain_nfs': ('plain', 45, 80, 2048,
                  'BMCNet_plain x4, NFS LR 45x80 -> 180x320, pretrain/BMCNet_plain_nfs_x4.pth (BASELINE config 2)'),
    'bmcnet_nfs': ('full', 45, 80, 2048,
                   'BMCNet x4, NFS LR 45x80 -> 180x320, surrogate weights (checkpoint not shipped)'),
    'bmcnet_eventzoom': ('full', 31, 56, 1024,
                         'BMCNet x4, EventZoom LR 31x56 -> 124x224, surrogate weights (checkpoint not shipped)'),
}
# tiles per conv job = ceil(B * R / 256), R = roundup((H+2)(W+2), 128): multiples of 19 images give
# 19 x 3968 / 256 = 294.5 ~ 2 x 148 tiles per job (whole waves of the 148 SMs); larger batches amortise
# the per-launch prologue (measured: plain 20.0k / 22.3k / 23.5k / 24.6k / 24.7k frames/s at B = 19 / 38 / 57 / 76 / 95,
# BMCNet 5.86k / 6.11k / 6.21k at B = 38 / 57 / 76)
DEFAULT_BATCH = {'plain_nfs': 95, 'bmcnet_nfs': 76, 'bmcnet_eventzoom': 156}
CONV_MAC_PER_PX = 147456          # 3x3 128->128
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (profiles/r0?_ncu_full_*.txt):
# conv_slab2_tc, 2 jobs, 45x80: B=95: 193.4 + 143.0 MB (algorithmic 193.0 read + 193.0 written; part of the output is
# still dirty in L2 when the kernel ends), B=57: 116.2 + 70.5 MB; encoders: bytes per event at 1e8 events
# (12.04 / 16.04 for 12 / 16 algorithmic)
NCU_CONV_TRAFFIC = {('plain', 95, 45, 80): (337.9e6, 'profiles/r02_ncu_full_slab2_plain3x3.txt'),     # 193.5 MB read + 144.5 MB written
                    ('plain', 57, 45, 80): (186.7e6, 'profiles/r01_B57_B38/'),
                    ('full', 76, 45, 80): (569.8e6, 'profiles/r01_ncu_full_slab2_bmcnet3x3.txt')}    # 4 jobs: 309.2 MB read + 260.5 MB written
NCU_ENC_BYTES_PER_EVENT, NCU_VOX_BYTES_PER_EVENT = 12.045, 16.035
FLOP_PER_PX = {'plain': 9721856, 'full': 41574912}      # SURVEY 8d / BASELINE.md section 3
SUSTAINED_SECONDS = 2.0


def load_state(model_kind):
    import torch
    from oracle import bmcnet_fp32 as O
    ck = os.path.join(ROOT, 'oracle', '_ref', 'BMCNet_plain_nfs_x4.pth')
    plain_sd = torch.load(ck, map_location='cpu') if os.path.exists(ck) else None
    if model_kind == 'plain':
        if plain_sd is not None:
            return plain_sd, 'shipped checkpoint'
        return O.surrogate_state_dict(plain=True), 'seeded surrogate (checkpoint not staged)'
    return O.surrogate_state_dict(plain=False, transplant=plain_sd), 'seeded surrogate + plain transplant'


def synth_stream(n_steps, batch, n_win, h, w, seed, device):
    """[n_steps, 3, batch*2*n_win] float32: xs, ys, ps of the two windows of every sequence."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    n = batch * 2 * n_win
    ev = torch.empty(n_steps, 3, n, device=device)
    ev[:, 0] = torch.randint(0, w, (n_steps, n), device=device, generator=g).float()
    ev[:, 1] = torch.randint(0, h, (n_steps, n), device=device, generator=g).float()
    ev[:, 2] = torch.randint(0, 2, (n_steps, n), device=device, generator=g).float() * 2 - 1
    return ev


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.first = [], None, 0
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def mark(self):
        """A timed region starts now: only samples that arrive from here on are reported.  (nvidia-smi needs
        100-300 ms to deliver its first sample, so the process is started before the warm-up steps.)"""
        self.first = len(self.rows)

    def report(self):
        """Summary of the samples since the last mark(); the sampler keeps running."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        rows, window = self.rows[self.first:], 'timed region'
        if not any(r and r[0].isdigit() for r in rows):       # region shorter than the sampling period
            rows, window = self.rows, 'warm-up + timed region (the timed region was shorter than one sampling period)'
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == 'Active'})
        busy = [v for v in sm if mx and v > 0.5 * mx[0]] or sm
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': reasons, 'samples': len(sm), 'window': window}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


E2E_RUNS = 3       # timed end-to-end regions of K steps each; the median is reported (see the e2e arm)
CPU_BATCH = 8      # sequences per CPU step: a bounded sample of the GPU arm's batch (batching helps oneDNN: 8.8 -> 13.2 frames/s on 8 cores)


def cpu_reference(model_kind, h, w, n_win, steps, warmup, batch=CPU_BATCH):
    """The reference's CPU path: its own modules + its own events_to_channels from the staged tree (baseline/_ref;
    kind 'reference'), else the oracle port (numpy encoder + functional fp32 forward; kind 'port').
    Returns (frames/s, ms/step, threads, kind)."""
    import numpy as np
    import torch
    from oracle import reference_tree as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, _ = load_state(model_kind)
    n_state = 2 if model_kind == 'plain' else 4
    st = [torch.zeros(batch, 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(batch, 32, h, w)]
    rng = np.random.default_rng(0)
    if R.available():
        kind = 'reference'
        BMCNet, BMCNet_plain = R.models()
        enc = R.encodings()
        model = (BMCNet_plain if model_kind == 'plain' else BMCNet)(4, 128, 5)      # infer_BMCNet.py:111
        model.load_state_dict(sd)
        model.eval()
        fwd = lambda x, *s: model(x, *s)
        encode = lambda xs, ys, ps: enc.events_to_channels(torch.from_numpy(xs), torch.from_numpy(ys), torch.from_numpy(ps),
                                                           sensor_size=(h, w))
    else:
        kind = 'port'
        from oracle import bmcnet_fp32 as O
        from oracle import encodings_np as E
        f = O.bmcnet_plain_forward if model_kind == 'plain' else O.bmcnet_forward
        fwd = lambda x, *s: f(sd, x, *s)
        encode = lambda xs, ys, ps: torch.from_numpy(E.events_to_channels(xs, ys, ps, sensor_size=(h, w)))

    def one(init):
        seqs = []
        for _ in range(batch):
            frames = []
            for _ in range(2):
                xs = rng.integers(0, w, n_win).astype(np.float32)
                ys = rng.integers(0, h, n_win).astype(np.float32)
                ps = rng.choice([-1.0, 1.0], n_win).astype(np.float32)
                frames.append(encode(xs, ys, ps))
            seqs.append(torch.stack(frames, 0))
        x = torch.stack(seqs, 0).transpose(1, 2)                 # [B,2,T,H,W] view, as infer_BMCNet.py:50
        with torch.no_grad():
            return list(fwd(x, *st, init))

    init = True
    for _ in range(warmup):
        st = one(init)
        init = False
    t0 = time.perf_counter()
    for _ in range(steps):
        st = one(init)
        init = False
    dt = time.perf_counter() - t0
    return steps * batch / dt, dt / steps * 1e3, cores, kind


def cpu_sample_text(kind, steps, warmup, n_win, cores):
    what = ('the reference\'s own modules + events_to_channels (baseline/_ref, unmodified)' if kind == 'reference'
            else 'numpy oracle encoder + functional fp32 PyTorch restatement of the reference forward')
    return ('%d recurrent steps of %d sequences after %d warm-up steps (a bounded sample of the GPU arm\'s batch), '
            '%d-event windows: %s, fp32 on %d threads' % (steps, CPU_BATCH, warmup, n_win, what, cores))


# ---------------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    pass


def run_workload(cx, name, batch, Ksteps, Wsteps, sustained=True):
    """The device-resident and the end-to-end timed regions of one workload on every rank; returns rank 0's
    numbers (max over ranks).  Leaves the model in cx.models[name]."""
    import torch
    import torch.distributed as dist
    from bmcnet_esr_b200.dataloader import encodings as G
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    model_kind, h, w, n_win, desc = WORKLOADS[name]
    dev, world, rank = cx.dev, cx.world, cx.rank
    B = batch
    sd, weights_desc = load_state(model_kind)
    model = (BMCNet_plain if model_kind == 'plain' else BMCNet)(4, 128, 5)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    cx.models[name] = model
    n_ev = B * 2 * n_win
    offsets = torch.arange(0, n_ev + 1, n_win, dtype=torch.int64, device=dev)
    total_steps = Wsteps + Ksteps
    stream = synth_stream(total_steps, B, n_win, h, w, 1234 + rank, dev)      # distinct events every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxed(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ------------------------------------------------------------------ device-resident arm
    def step_resident(k, reset):
        ev = stream[k % total_steps]
        cnt = G.events_to_channels_windows(ev[0], ev[1], ev[2], offsets, sensor_size=(h, w))   # [2B,2,H,W]
        x = cnt.view(B, 2, 2, h, w).transpose(1, 2)              # [B,2,T,H,W] view, as infer_BMCNet.py:50
        return model.step(x, reset=reset)

    for k in range(Wsteps):
        step_resident(k, k == 0)
    barrier()
    if cx.sampler:
        cx.sampler.mark()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(Ksteps):
        step_resident(Wsteps + k, False)
    t1.record()
    barrier()
    ms_total = maxed(t0.elapsed_time(t1))
    clocks = cx.sampler.report() if cx.sampler else None
    launches = (model._engine.launches_per_step + 3) * Ksteps        # graph nodes + encode, pack, emit
    frames = B * Ksteps * world
    res = {'workload': desc, 'batch_per_gpu': B, 'lr_hw': [h, w], 'events_per_window': n_win, 'weights': weights_desc,
           'value': frames / (ms_total * 1e-3), 'unit': 'frames/s', 'steps': Ksteps, 'ms_per_step': ms_total / Ksteps,
           'gpu_launches': launches, 'clocks': clocks,
           'model_gflop_per_frame': FLOP_PER_PX[model_kind] * h * w / 1e9}
    res['model_tflops'] = FLOP_PER_PX[model_kind] * h * w * res['value'] / 1e12

    # ------------------------------------------------------------------ end-to-end arm (timed right after the device-resident
    # region, i.e. in the same clock regime; the >= 2 s sustained run, which ends power-limited, comes last)
    host_ev = torch.empty(total_steps, 3, n_ev).pin_memory()           # step k reads row k % total_steps
    host_ev.copy_(stream)
    host_pred = [torch.empty(B, 2, 4 * h, 4 * w).pin_memory() for _ in range(2)]
    dev_ev = [torch.empty(3, n_ev, device=dev) for _ in range(2)]
    n_state = 2 if model_kind == 'plain' else 4
    # Copies ride a second stream so the H2D of step k+1 and the D2H of step k overlap the compute of the
    # neighbouring steps (what a serving loop does); every byte still moves inside the timed region and both
    # streams are drained before the closing event.
    main_s = torch.cuda.current_stream()
    copy_s = torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]       # events[k] landed in dev_ev[k & 1]
    ev_used = [torch.cuda.Event() for _ in range(2)]     # compute is done reading dev_ev[k & 1]

    def fresh_state():
        return [torch.zeros(B, 128, h, w, device=dev) for _ in range(n_state - 1)] + [torch.zeros(B, 32, h, w, device=dev)]

    def upload(k):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_used[k & 1])                                          # buffer free again
            dev_ev[k & 1].copy_(host_ev[k % total_steps], non_blocking=True)           # H2D, pinned
            ev_in[k & 1].record(copy_s)

    def step_e2e(k, st, init):
        upload(k + 1)
        main_s.wait_event(ev_in[k & 1])
        d = dev_ev[k & 1]
        cnt = G.events_to_channels_windows(d[0], d[1], d[2], offsets, sensor_size=(h, w))
        ev_used[k & 1].record(main_s)
        x = cnt.view(B, 2, 2, h, w).transpose(1, 2)
        with torch.no_grad():                                                        # infer_BMCNet.py:144 (@torch.no_grad())
            st = list(model(x, *st, init))                                           # the reference-facing call
        done = torch.cuda.Event()
        done.record(main_s)
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(done)
            host_pred[k & 1].copy_(st[-1], non_blocking=True)                        # D2H of the prediction
            st[-1].record_stream(copy_s)
        return st

    for e in ev_used:
        e.record(main_s)
    st = fresh_state()
    upload(0)
    for k in range(Wsteps):
        st = step_e2e(k, st, k == 0)
    # The end-to-end region depends on the host (PCIe, the Python thread): one run in ~6 on the shared boxes came out
    # 1.5-2x slow with the device-resident number unchanged.  It is therefore timed E2E_RUNS times, each run EXACTLY
    # K steps bracketed like the main region, and the MEDIAN is reported; all runs are listed in the JSON line.
    e2e_runs = []
    for run in range(E2E_RUNS):
        copy_s.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(Ksteps):
            st = step_e2e(Wsteps + run * Ksteps + k, st, False)    # K uploads (of the next step's events) + K read-backs inside
        main_s.wait_stream(copy_s)                                                   # the last read-back is inside the timed region
        e1.record()
        barrier()
        e2e_runs.append(maxed(e0.elapsed_time(e1)))
    ms_e2e = sorted(e2e_runs)[len(e2e_runs) // 2]
    res['e2e'] = {'value': frames / (ms_e2e * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms_e2e / Ksteps,
                  'h2d_bytes_per_step': 3 * n_ev * 4, 'd2h_bytes_per_step': B * 2 * 16 * h * w * 4,
                  'runs_frames_per_s': [frames / (t * 1e-3) for t in e2e_runs], 'reported': 'median of %d runs of K steps' % E2E_RUNS}
    # ------------------------------------------------------------------ sustained: the same loop for >= 2 s
    if sustained:
        n_sus = max(Ksteps, int(SUSTAINED_SECONDS * 1e3 / (ms_total / Ksteps)) + 1)
        if world > 1:                              # every rank must run the same number of steps
            t = torch.tensor([n_sus], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_sus = int(t.item())
        barrier()
        if cx.sampler:
            cx.sampler.mark()
        t0.record()
        for k in range(n_sus):
            step_resident(k, False)
        t1.record()
        barrier()
        ms_sus = maxed(t0.elapsed_time(t1))
        res['value_sustained'] = B * n_sus * world / (ms_sus * 1e-3)
        res['sustained'] = {'steps': n_sus, 'seconds': ms_sus * 1e-3, 'ms_per_step': ms_sus / n_sus,
                            'clocks': cx.sampler.report() if cx.sampler else None}

    res['arena_mb'] = model._engine.workspace.numel() / 1e6
    res['weights_mb'] = model._engine.weight_buf.numel() / 1e6
    del host_ev, host_pred, dev_ev, stream, st
    return res


TRAIN_LR = 1e-5


def train_targets(xs, rank):
    """Synthetic targets of the training workload: the x4 bilinear interpolation of each step's second count frame -- the
    base the model adds its learned residual to (BMCNet.py:119) -- plus N(0, 0.05^2) noise, i.e. a fine-tuning regime with
    small residual errors.  (Pure-noise Poisson targets on the surrogate weights make BPTT spike in fp32 autograd as well:
    max |grad| 0.6 -> 94 within three iterations of the reference's optimiser on one rank's sequence)"""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5 + rank)
    out = []
    for x in xs:
        base = F.interpolate(x[:, :, 1].float(), scale_factor=4, mode='bilinear', align_corners=False)
        out.append(base + 0.05 * torch.randn(base.shape, generator=g).to(base.device))
    return out


def run_train(cx, batch=2, seq_steps=8, h=45, w=80, iters=3, warm=1):
    """BASELINE config 5: the BMCNet x4 training iteration of train.py:202-237 (8 recurrent steps with BPTT, summed
    MSE, Adam(amsgrad); config/train_nfs.yml: batch 2, NFS LR 45x80), data-parallel: every rank runs its own
    sequences through the kernels, ONE all-reduce averages the 2,731,680 alias-deduplicated gradients, fused Adam."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models._train import FusedAdamAMSGrad, GraphedIteration, allreduce_gradients
    from oracle.make_golden import synth_counts
    dev, world, rank = cx.dev, cx.world, cx.rank
    sd, weights_desc = load_state('full')
    model = BMCNet(4, 128, 5)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    # Adam(amsgrad), wd 1e-5 as config/train_nfs.yml:28-34, but lr 1e-5 instead of 1e-4: the BMCNet weights here are the
    # surrogate set (the trained checkpoint is not in the mount) and sit at the edge of recurrent stability -- at 1e-4 the
    # fp32 reference arithmetic itself spikes on one sequence in eight (max |grad| 1.1 -> 25 -> 0.5, loss 0.63 -> 0.73 ->
    # 0.34) and the 16-bit path does not recover from that spike.  Timing does not depend on lr.
    opt = FusedAdamAMSGrad(model.parameters(), lr=TRAIN_LR)
    xs = [synth_counts(batch, h, w, 3000 + 17 * rank + s).to(dev) for s in range(seq_steps)]
    gts = train_targets(xs, rank)
    n_red = 0

    def iteration():
        nonlocal n_red
        opt.zero_grad()
        st = [torch.zeros(batch, 128, h, w, device=dev) for _ in range(3)] + [torch.zeros(batch, 32, h, w, device=dev)]
        loss, init = 0, True
        for x, gt in zip(xs, gts):
            st = list(model(x, *st, init))
            init = False
            loss = loss + F.mse_loss(st[-1], gt)
        loss.backward()
        if world > 1:
            n_red = allreduce_gradients(opt)
        opt.step()
        return loss.detach()

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            out = fn()
        t1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / n, out

    # (1) the iteration issued op by op from Python, as the reference's loop would run it on the drop-in module
    for _ in range(warm):
        iteration()
    ms_eager, loss = timed(iteration, iters)
    # (2) the same iteration with zero_grad .. backward replayed as ONE CUDA graph (GraphedIteration): the product path
    graphed = GraphedIteration(model, opt, xs, gts, warmup=1)
    n_red = opt.grad.numel() if world > 1 else 0
    graphed()
    ms, loss = timed(graphed, max(iters, 5))
    finite = bool(torch.isfinite(loss))
    del graphed, model, opt
    torch.cuda.empty_cache()
    return {'workload': 'BMCNet x4 training iteration (train.py:202-237): %d recurrent steps with BPTT, summed MSE, '
                        'Adam(amsgrad, lr %g), NFS LR %dx%d, batch %d per GPU, data-parallel; targets = x4 bilinear base of the '
                        'input counts + N(0, 0.05^2)' % (seq_steps, TRAIN_LR, h, w, batch),
            'weights': weights_desc, 'ms_per_iteration': ms, 'iterations_per_s': 1e3 / ms,
            'ms_per_iteration_eager': ms_eager, 'how': 'zero_grad + forward + backward replayed as one CUDA graph, then the '
            'all-reduce and the fused Adam launch (GraphedIteration); `ms_per_iteration_eager` = the same iteration issued '
            'op by op from Python (host-bound)',
            'value': batch * seq_steps * world / (ms * 1e-3), 'unit': 'training frames/s (forward + backward + update)',
            'iterations': iters, 'batch_per_gpu': batch, 'sequence_steps': seq_steps, 'loss_finite': finite,
            'allreduce': {'elements': n_red, 'bytes': n_red * 4, 'collective': 'one NCCL all-reduce of the flat '
                          'alias-deduplicated fp32 gradient buffer per iteration'} if world > 1 else None,
            'dtype': 'f16 activations / activation gradients under a static loss scale, fp32 weight gradients and Adam',
            'kernels': 'every convolution forward / dgrad (bmc_conv_gemm) and wgrad (bmc_conv_wgrad) on tcgen05; LayerNorm, '
                       '128x128 attention and layout glue are PyTorch ops'}


def latency_b1(cx, model_kind, h, w, iters=60):
    """Batch-1 latency exactly as infer_BMCNet.py:54-68 measures it: a CUDA-event pair around the forward call of one
    recording's recurrent loop, synchronised every frame (so launch latency and the host side of `forward` count)."""
    import torch
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    from oracle.make_golden import synth_counts
    dev = cx.dev
    sd, _ = load_state(model_kind)
    model = (BMCNet_plain if model_kind == 'plain' else BMCNet)(4, 128, 5)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    return _latency_loop(model, model_kind, h, w, iters, dev, synth_counts), model


def _latency_loop(model, model_kind, h, w, iters, dev, synth_counts, batch=1):
    import torch
    n_state = 2 if model_kind == 'plain' else 4
    xs = [synth_counts(batch, h, w, 50 + i).transpose(1, 2).contiguous().to(dev).transpose(1, 2) for i in range(8)]
    st = [torch.zeros(batch, 128, h, w, device=dev) for _ in range(n_state - 1)] + [torch.zeros(batch, 32, h, w, device=dev)]
    starter, ender = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    with torch.no_grad():
        for i in range(iters + 5):
            starter.record()
            st = list(model(xs[i % 8], *st, i == 0))
            ender.record()
            torch.cuda.synchronize()
            if i >= 5:
                times.append(starter.elapsed_time(ender))
    times.sort()
    return {'ms_per_frame_median': times[len(times) // 2], 'ms_per_frame_min': times[0], 'frames': len(times)}


def eager_b200(cx, model_kind, h, w, batch, iters_b1=30, iters_batched=4):
    """The unmodified reference modules on this B200 in eager mode (stock settings), batch 1 and batched."""
    import torch
    from oracle import reference_tree as R
    from oracle.make_golden import synth_counts
    if not R.available():
        return {'unavailable': 'baseline/_ref not staged'}
    BMCNet, BMCNet_plain = R.models()
    sd, _ = load_state(model_kind)
    model = (BMCNet_plain if model_kind == 'plain' else BMCNet)(4, 128, 5)
    model.load_state_dict(sd)
    model = model.to(cx.dev).eval()
    out = {'what': 'reference %s (baseline/_ref, unmodified) .to(cuda).eval(), fp32 eager, torch %s' % (
        'BMCNet_plain' if model_kind == 'plain' else 'BMCNet', torch.__version__)}
    out['b1'] = _latency_loop(model, model_kind, h, w, iters_b1, cx.dev, synth_counts)
    try:
        lb = _latency_loop(model, model_kind, h, w, iters_batched, cx.dev, synth_counts, batch=batch)
        out['batched'] = {'batch': batch, 'ms_per_step_median': lb['ms_per_frame_median'],
                          'frames_per_s': batch / (lb['ms_per_frame_median'] * 1e-3)}
    except torch.cuda.OutOfMemoryError:
        out['batched'] = {'batch': batch, 'unavailable': 'out of memory in eager fp32'}
    del model
    torch.cuda.empty_cache()
    return out


def conv_roofline(cx, model_kind, B, h, w, peaks):
    import torch
    from bmcnet_esr_b200 import _lib, kernels as K
    dev = cx.dev
    jobs = 2 if model_kind == 'plain' else 4         # ResidualBlock convs run 2 (plain) / 4 (BMCNet) jobs per launch
    # all jobs live in ONE tensor (like the model's activation arena) so the launch takes the
    # product path: the persistent slab kernel with a single TMA descriptor per box shape
    src = K.pack_nchw(torch.randn(jobs * B, 128, h, w, device=dev))
    rows_job = src.shape[0] // jobs
    wpk = K.pack_conv_weight(torch.randn(128, 128, 3, 3, device=dev) * 0.03, [(0, 128)])
    bias = torch.zeros(128, device=dev)
    jarr = (_lib.GemmJob * jobs)()
    outs = torch.empty_like(src)
    for j in range(jobs):
        jarr[j].n_seg = 1
        jarr[j].a[0] = src.data_ptr(); jarr[j].a_rows[0] = src.shape[0]; jarr[j].a_ch[0] = 128
        jarr[j].a_row_base[0] = j * rows_job
        jarr[j].w = wpk.data_ptr(); jarr[j].w_rows = 128; jarr[j].w_k = 1152
        jarr[j].bias = bias.data_ptr(); jarr[j].out_act16 = outs.data_ptr(); jarr[j].out_row_base = j * rows_job
        jarr[j].relu = 1
    launch = lambda: _lib.check(_lib.lib().bmc_conv_gemm(jarr, jobs, 128, 9, B, h, w, 0, _lib.stream_ptr()))
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    reps = 50
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                launch()
    g.replay()
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    g.replay()
    r1.record()
    torch.cuda.synchronize()
    conv_ms = r0.elapsed_time(r1) / reps
    conv_flops = 2.0 * CONV_MAC_PER_PX * h * w * B * jobs
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    peak = peaks.get('bf16_tflops', 1590.0)
    return {'kernel': 'conv_slab2_tc (3x3 128->128 implicit GEMM, %d jobs, B=%d)' % (jobs, B), 'bound': 'tensor',
            'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
            'traffic': NCU_CONV_TRAFFIC.get((model_kind, B, h, w), (None, None))[0],
            'traffic_unit': 'bytes per launch (ncu --set full, %s)' % NCU_CONV_TRAFFIC.get((model_kind, B, h, w), (None, 'no capture at this shape'))[1],
            'algorithmic_bytes': 2.0 * jobs * B * (((h + 2) * (w + 2) + 127) // 128 * 128) * 128 * 2,   # input read + output written
            'us_per_launch': conv_ms * 1e3,
            'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)' if peaks else 'fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)'}


def encoder_rooflines(cx, h, w, peaks):
    import torch
    from bmcnet_esr_b200.dataloader import encodings as G
    dev = cx.dev
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_big = 400_000_000           # 4.8 GB of events: the ~30 us of launch / memset / finalize per call are < 4 % of it
    xs = torch.rand(n_big, device=dev) * w
    ys = torch.rand(n_big, device=dev) * h
    ps = (torch.rand(n_big, device=dev) < 0.5).float() * 2 - 1
    for _ in range(3):
        G.events_to_channels(xs, ys, ps, sensor_size=(h, w))
    torch.cuda.synchronize()
    r0.record()
    for _ in range(5):
        G.events_to_channels(xs, ys, ps, sensor_size=(h, w))
    r1.record()
    torch.cuda.synchronize()
    enc_ms = r0.elapsed_time(r1) / 5
    enc_gbs = (12.0 * n_big + 2 * h * w * 4) / (enc_ms * 1e-3) / 1e9
    hbm = peaks.get('hbm_gbs', 6650.0)
    # time-interpolated voxels (events_to_voxel, 5 bins): 16 B/event (BASELINE metric "voxel encoding Mevents/s")
    n_vox = 400_000_000
    ts = torch.sort(torch.rand(n_vox, device=dev))[0]
    for _ in range(2):
        G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w))
    torch.cuda.synchronize()
    r0.record()
    for _ in range(3):
        G.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(h, w))
    r1.record()
    torch.cuda.synchronize()
    vox_ms = r0.elapsed_time(r1) / 3
    vox_gbs = (16.0 * n_vox + 5 * h * w * 4) / (vox_ms * 1e-3) / 1e9
    del xs, ys, ps, ts
    torch.cuda.empty_cache()
    enc = {'kernel': 'scatter_kernel<ChannelsOp> (events_to_channels, %.0e events, %dx%d)' % (n_big, h, w),
           'bound': 'hbm', 'achieved': enc_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': enc_gbs / hbm,
           'traffic': NCU_ENC_BYTES_PER_EVENT * n_big if (h, w) == (45, 80) else None,
           'note': 'a read-only stream: it can exceed the peak, which is a copy (read + write) bandwidth',
           'mevents_per_s': n_big / (enc_ms * 1e-3) / 1e6}
    vox = {'kernel': 'scatter_kernel<VoxelOp> (events_to_voxel, 5 bins, %.0e events, %dx%d)' % (n_vox, h, w),
           'bound': 'hbm', 'achieved': vox_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': vox_gbs / hbm,
           'traffic': NCU_VOX_BYTES_PER_EVENT * n_vox if (h, w) == (45, 80) else None,
           'mevents_per_s': n_vox / (vox_ms * 1e-3) / 1e6}
    return enc, vox


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', default='plain_nfs', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0,
                    help='independent sequences per GPU, stepped in lockstep (default: per workload, chosen so the '
                         '256-row conv tiles fill whole waves of the 148 SMs)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-steps', type=int, default=6)
    ap.add_argument('--only-headline', action='store_true',
                    help='skip the extra workloads / latency / eager legs (kernel work loops, profiling)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    model_kind, h, w, n_win, desc = WORKLOADS[args.workload]
    if args.batch <= 0:
        args.batch = DEFAULT_BATCH[args.workload]
    config = {'workload': desc, 'batch_per_gpu': args.batch, 'lr_hw': [h, w], 'events_per_window': n_win,
              'windows_per_step_per_sequence': 2, 'sharding': 'independent sequences per GPU, no collective'}

    if args.impl == 'reference':
        if rank != 0:
            return
        # every step the driver asks for, each on a bounded sample (CPU_BATCH sequences) of the GPU arm's batch
        fps, ms, cores, kind = cpu_reference(model_kind, h, w, n_win, args.steps, args.warmup)
        _, weights_desc = load_state(model_kind)
        print(json.dumps({
            'impl': 'reference', 'metric': 'x4_sr_event_frames_per_sec', 'value': fps, 'unit': 'frames/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(config, weights=weights_desc),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                             'sample': cpu_sample_text(kind, args.steps, args.warmup, n_win, cores)},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    from bmcnet_esr_b200 import _lib

    torch.cuda.set_device(local_rank)
    cx = Ctx()
    cx.dev = torch.device('cuda', local_rank)
    cx.world, cx.rank, cx.models = world, rank, {}
    if world > 1:
        dist.init_process_group('nccl', device_id=cx.dev)
    cx.sampler = ClockSampler(local_rank) if rank == 0 else None

    head = run_workload(cx, args.workload, args.batch, args.steps, args.warmup)
    extra = {}
    if not args.only_headline:
        del cx.models[args.workload]
        torch.cuda.empty_cache()
        for name in WORKLOADS:
            if name != args.workload:
                extra[name] = run_workload(cx, name, DEFAULT_BATCH[name], args.steps, args.warmup)
                del cx.models[name]
                torch.cuda.empty_cache()
        extra['bmcnet_train_nfs'] = run_train(cx)
        if world == 1:
            # the reference's batch (2) leaves the iteration host-bound (Python autograd tape); a larger batch shows
            # what the kernels sustain
            try:
                big = run_train(cx, batch=8, iters=2)
                extra['bmcnet_train_nfs']['batch_8'] = {k: big[k] for k in ('value', 'unit', 'ms_per_iteration', 'batch_per_gpu', 'loss_finite')}
            except torch.cuda.OutOfMemoryError:
                extra['bmcnet_train_nfs']['batch_8'] = {'unavailable': 'out of memory'}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = head.pop('clocks')
    if cx.sampler:
        cx.sampler.stop()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roofline = conv_roofline(cx, model_kind, args.batch, h, w, peaks)
    enc_roof, vox_roof = encoder_rooflines(cx, h, w, peaks)

    lat, eager = {}, {}
    if not args.only_headline and world == 1:
        for kind_, name_ in (('plain', 'BMCNet_plain'), ('full', 'BMCNet')):
            r, m = latency_b1(cx, kind_, 45, 80)
            r['launches_per_frame'] = m._engine.launches_per_step + 2 + (2 if kind_ == 'plain' else 4)   # graph nodes + pack/emit + state pack/unpack
            lat[name_ + '_45x80'] = r
            del m
        torch.cuda.empty_cache()
        eager['BMCNet_plain_45x80'] = eager_b200(cx, 'plain', 45, 80, DEFAULT_BATCH['plain_nfs'])
        eager['BMCNet_45x80'] = eager_b200(cx, 'full', 45, 80, DEFAULT_BATCH['bmcnet_nfs'])

    # ------------------------------------------------------------------ CPU baseline (bounded sample)
    cpu_fps, cpu_ms, cores, cpu_kind = cpu_reference(model_kind, h, w, n_win, args.cpu_steps, 2)

    line = {
        'metric': 'x4_sr_event_frames_per_sec', 'value': head['value'], 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': _lib.lib().bmc_act_dtype().decode(), 'data': 'synthetic',
        'config': dict(config, weights=head['weights'],
                       l2='every step streams distinct events; weights (%.1f MB) stay L2-resident by design; '
                          'activation arena %.0f MB' % (head['weights_mb'], head['arena_mb'])),
        'e2e': head['e2e'],
        'value_sustained': head.get('value_sustained'), 'sustained': head.get('sustained'),
        'gpu_launches': head['gpu_launches'],
        'clocks': clocks,
        'roofline': roofline,
        'roofline_encoder': enc_roof,
        'roofline_voxel': vox_roof,
        'model_gflop_per_frame': head['model_gflop_per_frame'],
        'model_tflops': head['model_tflops'],
        'cpu_baseline': {'value': cpu_fps, 'unit': 'frames/s', 'cores': cores, 'kind': cpu_kind,
                         'sample': cpu_sample_text(cpu_kind, args.cpu_steps, 2, n_win, cores)},
    }
    if extra:
        line['workloads'] = extra
    if lat:
        line['latency_b1'] = dict(lat, how='CUDA events around forward(), synchronised per frame (infer_BMCNet.py:54-68), '
                                           'median of 60 frames after 5 warm-up frames, B=1')
    if eager:
        line['eager_b200'] = eager
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
