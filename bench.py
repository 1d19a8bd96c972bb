"""Headline benchmark: x4 SR event-frames/s of the BMCNet hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload plain_nfs|bmcnet_nfs|bmcnet_eventzoom]
                    [--batch B] [--impl reference]

One STEP = one recurrent step of the hot path for a batch of B independent synthetic sequences on
each GPU: encode the step's two event windows per sequence into per-polarity count frames
(`events_to_channels`, the encoder on the live path) and run one `forward` of the model
(x4 prediction for every sequence).  value = event-frames/s summed over all GPUs.

  value   device-resident: events already in HBM, recurrent state kept in the arena
          (bmc_model_step), prediction written to HBM.
  e2e     through the reference-facing API (`model(x, h, o, init)` + `events_to_channels_windows`)
          with the step's events copied from pinned host memory and the prediction read back to
          the host inside the timed region, every step; timed 3 x K steps, the median run is reported
          (all three are listed: the host side of a shared box is noisy).
  roofline  the dominant kernel (3x3 128->128 implicit-GEMM conv, tcgen05: conv_slab2_tc) timed live,
          back to back, at this workload's shape; algorithmic FLOPs = 2*147456 MAC per real LR pixel;
          `traffic` = DRAM bytes per launch of the committed ncu --set full capture (profiles/).
  cpu_baseline / --impl reference  the CPU oracle (fp32 PyTorch restatement of the reference
          forward + numpy encoder) on all host cores, on a bounded sample (8 sequences per step).

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
CONV_MAC_PER_PX = 147456          # 3x3 128->128
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (profiles/r01_ncu_full_*.txt):
# conv_slab2_tc, 2 jobs, 45x80: B=95: 193.4 + 143.0 MB (algorithmic 193.0 read + 193.0 written; part of the output is
# still dirty in L2 when the kernel ends), B=57: 116.2 + 70.5 MB; encoders: bytes per event at 1e8 events
# (12.04 / 16.04 for 12 / 16 algorithmic)
NCU_CONV_TRAFFIC = {('plain', 95, 45, 80): 336.4e6, ('plain', 57, 45, 80): 186.7e6,
                    ('full', 76, 45, 80): 569.8e6}       # 4 jobs, B=76: 309.2 MB read + 260.5 MB written
NCU_ENC_BYTES_PER_EVENT, NCU_VOX_BYTES_PER_EVENT = 12.045, 16.035
FLOP_PER_PX = {'plain': 9721856, 'full': 41574912}      # SURVEY 8d / BASELINE.md section 3


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
        self.rows, self.proc = [], None
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
        """The timed region starts now: only samples that arrive from here on are reported.  (nvidia-smi needs
        100-300 ms to deliver its first sample, so the process is started before the warm-up steps.)"""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows, window = self.rows[getattr(self, 'first', 0):], 'timed region'
        if not any(r and r[0].isdigit() for r in rows):       # region shorter than the sampling period
            rows, window = self.rows, 'warm-up + timed region (the timed region was shorter than one sampling period)'
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == 'Active'})
        busy = [v for v in sm if mx and v > 0.5 * mx[0]] or sm
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': reasons, 'samples': len(sm), 'window': window}


E2E_RUNS = 3       # timed end-to-end regions of K steps each; the median is reported (see the e2e arm)
CPU_BATCH = 8      # sequences per CPU step: a bounded sample of the GPU arm's batch (batching helps oneDNN: 8.8 -> 13.2 frames/s on 8 cores)


def cpu_reference(model_kind, h, w, n_win, steps, warmup, batch=CPU_BATCH):
    """The CPU path of the reference (oracle port): numpy encoder + fp32 PyTorch forward on `batch` sequences."""
    import numpy as np
    import torch
    from oracle import bmcnet_fp32 as O
    from oracle import encodings_np as E
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, _ = load_state(model_kind)
    fwd = O.bmcnet_plain_forward if model_kind == 'plain' else O.bmcnet_forward
    n_state = 2 if model_kind == 'plain' else 4
    st = [torch.zeros(batch, 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(batch, 32, h, w)]
    rng = np.random.default_rng(0)

    def one(init):
        seqs = []
        for _ in range(batch):
            frames = []
            for _ in range(2):
                xs = rng.integers(0, w, n_win).astype(np.float32)
                ys = rng.integers(0, h, n_win).astype(np.float32)
                ps = rng.choice([-1.0, 1.0], n_win).astype(np.float32)
                frames.append(torch.from_numpy(E.events_to_channels(xs, ys, ps, sensor_size=(h, w))))
            seqs.append(torch.stack(frames, 0))
        x = torch.stack(seqs, 0).transpose(1, 2)                 # [B,2,T,H,W] view, as infer_BMCNet.py:50
        return list(fwd(sd, x, *st, init))

    init = True
    for _ in range(warmup):
        st = one(init)
        init = False
    t0 = time.perf_counter()
    for _ in range(steps):
        st = one(init)
        init = False
    dt = time.perf_counter() - t0
    return steps * batch / dt, dt / steps * 1e3, cores


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    model_kind, h, w, n_win, desc = WORKLOADS[args.workload]
    if args.batch <= 0:
        # tiles per conv job = ceil(B * R / 256), R = roundup((H+2)(W+2), 128): multiples of 19 images give
        # 19 x 3968 / 256 = 294.5 ~ 2 x 148 tiles per job (whole waves of the 148 SMs); larger batches amortise
        # the per-launch prologue (measured: plain 20.0k / 22.3k / 23.5k / 24.6k / 24.7k frames/s at B = 19 / 38 / 57 / 76 / 95,
        # BMCNet 5.86k / 6.11k / 6.21k at B = 38 / 57 / 76)
        args.batch = {'plain_nfs': 95, 'bmcnet_nfs': 76, 'bmcnet_eventzoom': 156}[args.workload]
    config = {'workload': desc, 'batch_per_gpu': args.batch, 'lr_hw': [h, w], 'events_per_window': n_win,
              'windows_per_step_per_sequence': 2, 'sharding': 'independent sequences per GPU, no collective'}

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = min(args.steps, 12)
        fps, ms, cores = cpu_reference(model_kind, h, w, n_win, steps, min(args.warmup, 2))
        sample = '%d recurrent steps of %d sequences (a bounded sample of the GPU arm\'s batch), %d-event windows encoded ' \
                 'by the numpy oracle, fp32 PyTorch CPU forward, %d threads' % (steps, CPU_BATCH, n_win, cores)
        print(json.dumps({
            'impl': 'reference', 'metric': 'x4_sr_event_frames_per_sec', 'value': fps, 'unit': 'frames/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': min(args.warmup, 2), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(config, batch_per_gpu=CPU_BATCH),
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    from bmcnet_esr_b200 import _lib, kernels as K
    from bmcnet_esr_b200.dataloader import encodings as G
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, Ksteps, Wsteps = args.batch, args.steps, args.warmup

    sd, weights_desc = load_state(model_kind)
    model = (BMCNet_plain if model_kind == 'plain' else BMCNet)(4, 128, 5)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    config['weights'] = weights_desc

    n_ev = B * 2 * n_win
    offsets = torch.arange(0, n_ev + 1, n_win, dtype=torch.int64, device=dev)
    total_steps = Wsteps + Ksteps
    stream = synth_stream(total_steps, B, n_win, h, w, 1234 + rank, dev)      # distinct events every step
    preds = torch.empty(B, 2, 4 * h, 4 * w, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm
    def step_resident(k, reset):
        ev = stream[k]
        cnt = G.events_to_channels_windows(ev[0], ev[1], ev[2], offsets, sensor_size=(h, w))   # [2B,2,H,W]
        x = cnt.view(B, 2, 2, h, w).transpose(1, 2)              # [B,2,T,H,W] view, as infer_BMCNet.py:50
        return model.step(x, reset=reset)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for k in range(Wsteps):
        step_resident(k, k == 0)
    barrier()
    if sampler:
        sampler.mark()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(Ksteps):
        step_resident(Wsteps + k, False)
    t1.record()
    barrier()
    ms_total = torch.tensor([t0.elapsed_time(t1)], device=dev)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = ms_total.item()
    launches = (model._engine.launches_per_step + 3) * Ksteps        # graph nodes + encode, pack, emit

    # ------------------------------------------------------------------ end-to-end arm
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
    ev_out = [torch.cuda.Event() for _ in range(2)]      # host_pred[k & 1] has been read back

    def fresh_state():
        return [torch.zeros(B, 128, h, w, device=dev) for _ in range(n_state - 1)] + [torch.zeros(B, 32, h, w, device=dev)]

    def upload(k):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_used[k & 1])                                          # buffer free again
            dev_ev[k & 1].copy_(host_ev[k % total_steps], non_blocking=True)           # H2D, pinned
            ev_in[k & 1].record(copy_s)

    def step_e2e(k, st, init, last):
        if not last:
            upload(k + 1)
        main_s.wait_event(ev_in[k & 1])
        d = dev_ev[k & 1]
        cnt = G.events_to_channels_windows(d[0], d[1], d[2], offsets, sensor_size=(h, w))
        ev_used[k & 1].record(main_s)
        x = cnt.view(B, 2, 2, h, w).transpose(1, 2)
        st = list(model(x, *st, init))                                               # the reference-facing call
        done = torch.cuda.Event()
        done.record(main_s)
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(done)
            host_pred[k & 1].copy_(st[-1], non_blocking=True)                        # D2H of the prediction
            st[-1].record_stream(copy_s)
            ev_out[k & 1].record(copy_s)
        return st

    for e in ev_used:
        e.record(main_s)
    st = fresh_state()
    upload(0)
    for k in range(Wsteps):
        st = step_e2e(k, st, k == 0, False)
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
            st = step_e2e(Wsteps + run * Ksteps + k, st, False, False)    # K uploads (of the next step's events) + K read-backs inside
        main_s.wait_stream(copy_s)                                                   # the last read-back is inside the timed region
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_runs.append(t.item())
    ms_e2e = sorted(e2e_runs)[len(e2e_runs) // 2]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    jobs = 2 if model_kind == 'plain' else 4         # ResidualBlock convs run 2 (plain) / 4 (BMCNet) jobs per launch
    # all jobs live in ONE tensor (like the model's activation arena) so the launch takes the
    # product path: the persistent slab kernel with a single TMA descriptor per box shape
    src = K.pack_nchw(torch.randn(jobs * B, 128, h, w, device=dev))
    rows_job = src.shape[0] // jobs
    wpk = K.pack_conv_weight(torch.randn(128, 128, 3, 3, device=dev) * 0.03, [(0, 128)])
    bias = torch.zeros(128, device=dev)
    import ctypes as C
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
    roofline = {'kernel': 'conv_slab2_tc (3x3 128->128 implicit GEMM, %d jobs, B=%d)' % (jobs, B), 'bound': 'tensor',
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'traffic': NCU_CONV_TRAFFIC.get((model_kind, B, h, w)), 'traffic_unit': 'bytes per launch (ncu, profiles/r01_ncu_full_slab2_%s3x3.txt)' % ('plain' if model_kind == 'plain' else 'bmcnet'),
                'algorithmic_bytes': 2.0 * jobs * B * (((h + 2) * (w + 2) + 127) // 128 * 128) * 128 * 2,   # input read + output written
                'us_per_launch': conv_ms * 1e3,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)' if peaks else 'fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)'}

    # encoder kernel against HBM
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
    xv, yv, pv = xs[:n_vox].contiguous(), ys[:n_vox].contiguous(), ps[:n_vox].contiguous()
    for _ in range(2):
        G.events_to_voxel(xv, yv, ts, pv, 5, sensor_size=(h, w))
    torch.cuda.synchronize()
    r0.record()
    for _ in range(3):
        G.events_to_voxel(xv, yv, ts, pv, 5, sensor_size=(h, w))
    r1.record()
    torch.cuda.synchronize()
    vox_ms = r0.elapsed_time(r1) / 3
    vox_gbs = (16.0 * n_vox + 5 * h * w * 4) / (vox_ms * 1e-3) / 1e9
    del xs, ys, ps, xv, yv, pv, ts

    # ------------------------------------------------------------------ CPU baseline (bounded sample)
    cpu_fps, cpu_ms, cores = cpu_reference(model_kind, h, w, n_win, args.cpu_steps, 2)

    frames = B * Ksteps * world
    value = frames / (ms_total * 1e-3)
    line = {
        'metric': 'x4_sr_event_frames_per_sec', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': Ksteps,
        'warmup': Wsteps, 'ms_per_step': ms_total / Ksteps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': _lib.lib().bmc_act_dtype().decode(), 'data': 'synthetic',
        'config': dict(config, l2='every step streams distinct events; weights (%.1f MB) stay L2-resident by design; '
                                  'activation arena %.0f MB' % (model._engine.weight_buf.numel() / 1e6,
                                                                model._engine.workspace.numel() / 1e6)),
        'e2e': {'value': frames / (ms_e2e * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms_e2e / Ksteps,
                'h2d_bytes_per_step': 3 * n_ev * 4, 'd2h_bytes_per_step': B * 2 * 16 * h * w * 4,
                'runs_frames_per_s': [frames / (t * 1e-3) for t in e2e_runs], 'reported': 'median of %d runs of K steps' % E2E_RUNS},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': roofline,
        'roofline_encoder': {'kernel': 'scatter_kernel<ChannelsOp> (events_to_channels, %.0e events, %dx%d)' % (n_big, h, w),
                             'bound': 'hbm', 'achieved': enc_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': enc_gbs / hbm,
                             'traffic': NCU_ENC_BYTES_PER_EVENT * n_big if (h, w) == (45, 80) else None,
                             'note': 'a read-only stream: it can exceed the peak, which is a copy (read + write) bandwidth',
                             'mevents_per_s': n_big / (enc_ms * 1e-3) / 1e6},
        'roofline_voxel': {'kernel': 'scatter_kernel<VoxelOp> (events_to_voxel, 5 bins, %.0e events, %dx%d)' % (n_vox, h, w),
                           'bound': 'hbm', 'achieved': vox_gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': vox_gbs / hbm,
                           'traffic': NCU_VOX_BYTES_PER_EVENT * n_vox if (h, w) == (45, 80) else None,
                           'mevents_per_s': n_vox / (vox_ms * 1e-3) / 1e6},
        'model_gflop_per_frame': FLOP_PER_PX[model_kind] * h * w / 1e9,
        'model_tflops': FLOP_PER_PX[model_kind] * h * w * value / 1e12,
        'cpu_baseline': {'value': cpu_fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d recurrent steps of %d sequences after 2 warm-up steps: numpy oracle '
                                   'encoder + fp32 PyTorch CPU forward, %d threads' % (args.cpu_steps, CPU_BATCH, cores)},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
