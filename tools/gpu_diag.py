"""GPU bring-up diagnostics: each stage runs in its own process (a trap in one kernel must not
take the others down) and prints error statistics against torch / the CPU oracle.

    python tools/gpu_diag.py [stage ...]        # default: all stages
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ['enc', 'simt_conv', 'tc_conv', 'tc_conv_ln', 'tc_conv32', 'simt_att', 'tc_att', 'plain_simt',
          'plain_tc', 'full_simt', 'full_tc', 'step_tc', 'perf']


def rnd(*shape, scale=1.0, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def conv_case(impl, cin_list, n, taps, relu, residual, ln, b=2, h=13, w=21):
    """conv-gemm vs torch conv2d on bf16-rounded operands."""
    import torch
    import torch.nn.functional as F
    from bmcnet_esr_b200 import kernels as K
    dev = 'cuda'
    xs = [rnd(b, c, h, w, seed=10 + i).bfloat16().float() for i, c in enumerate(cin_list)]
    cin = sum(cin_list)
    k = 3 if taps == 9 else 1
    wt = rnd(n, cin, k, k, scale=1.0 / (cin * taps) ** 0.5, seed=3).bfloat16().float()
    bias = rnd(n, scale=0.1, seed=4)
    res = rnd(b, n, h, w, seed=5).bfloat16().float() if residual else None
    y = F.conv2d(torch.cat(xs, 1), wt, bias, padding=k // 2)
    lnp = None
    if ln:
        gam, bet = 1 + rnd(n, scale=0.1, seed=6), rnd(n, scale=0.1, seed=7)
        mu = y.mean(1, keepdim=True)
        var = (y - mu).pow(2).mean(1, keepdim=True)
        y = (y - mu) / (var + 1e-6).sqrt() * gam.view(1, -1, 1, 1) + bet.view(1, -1, 1, 1)
        lnp = (gam.to(dev), bet.to(dev), 1e-6)
    if relu:
        y = F.relu(y)
    if residual:
        y = y + res
    segs, first = [], 0
    for c in cin_list:
        segs.append((first, c))
        first += c
    wpk = K.pack_conv_weight(wt.to(dev), segs)
    srcs = [K.pack_nchw(x.to(dev)) for x in xs]
    rp = K.pack_nchw(res.to(dev)) if residual else None
    out, outf = K.conv_gemm(srcs, wpk, bias.to(dev), b, h, w, taps, n=n, relu=relu, residual=rp, ln=lnp,
                            impl=impl, out_f32=True)
    torch.cuda.synchronize()
    got = K.unpack_nchw(out, b, n, h, w).cpu()
    # fp32 side output: [rows, n] -> compare on interior through a bf16-free path
    R = K.rows_per_image(h, w)
    of = outf.view(b, R, n)[:, :(h + 2) * (w + 2)].view(b, h + 2, w + 2, n)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).cpu()
    halo = outf.view(b, R, n).clone()
    halo[:, :(h + 2) * (w + 2)].view(b, h + 2, w + 2, n)[:, 1:-1, 1:-1] = 0
    e32 = (of - y).abs().max().item()
    e16 = (got - y).abs().max().item()
    print('  impl=%d cin=%s n=%d taps=%d relu=%d res=%d ln=%d: max|err| fp32-out %.3e  bf16-out %.3e  (ref max %.3f) halo max %.1e'
          % (impl, cin_list, n, taps, relu, residual, ln, e32, e16, y.abs().max().item(), halo.abs().max().item()))
    return e32


def att_case(impl, b=2, h=13, w=21, n_split=3):
    import torch
    from bmcnet_esr_b200 import kernels as K
    c = rnd(b, 128, h, w, scale=0.5, seed=1).bfloat16().float()
    v = rnd(b, 128, h, w, scale=0.5, seed=2).bfloat16().float()
    scale = 128 ** -0.5
    att = torch.bmm(c.view(b, 128, -1), v.view(b, 128, -1).transpose(1, 2)) * scale
    pr = torch.softmax(att, -1)
    probs, partial = K.attention_weights(K.pack_nchw(c.cuda()), K.pack_nchw(v.cuda()), b, h, w, scale, n_split, impl)
    torch.cuda.synchronize()
    got_att = partial.sum(1).cpu()
    got_p = probs.view(b, 2, 128, 64).permute(0, 2, 1, 3).reshape(b, 128, 128).float().cpu()
    print('  impl=%d att max|err| %.3e (ref max %.2f)   softmax max|err| %.3e'
          % (impl, (got_att - att).abs().max().item(), att.abs().max().item(), (got_p - pr).abs().max().item()))
    out = K.apply_dynamic_weights(K.pack_nchw(v.cuda()), probs, b, h, w, impl=impl)
    torch.cuda.synchronize()
    ref = torch.bmm(pr.bfloat16().float(), v.view(b, 128, -1)).view(b, 128, h, w)
    got = K.unpack_nchw(out, b, 128, h, w).cpu()
    print('  impl=%d softmax(att)@v max|err| %.3e (ref max %.2f)' % (impl, (got - ref).abs().max().item(), ref.abs().max().item()))


def model_case(plain, simt, h=22, w=40, b=2, steps=3, use_step=False):
    import torch
    from oracle import bmcnet_fp32 as O
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    ck = os.path.join(ROOT, 'oracle', '_ref', 'BMCNet_plain_nfs_x4.pth')
    plain_sd = torch.load(ck, map_location='cpu') if os.path.exists(ck) else None
    if plain:
        sd = plain_sd if plain_sd is not None else O.surrogate_state_dict(plain=True)
        m = BMCNet_plain(4, 128, 5)
    else:
        sd = O.surrogate_state_dict(plain=False, transplant=plain_sd)
        m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m._engine.set_debug_simt(simt)
    n_state = 2 if plain else 4
    st_ref = [torch.zeros(b, 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w)]
    st = [t.cuda() for t in st_ref]
    fwd = O.bmcnet_plain_forward if plain else O.bmcnet_forward
    init = True
    for s in range(steps):
        g = torch.Generator().manual_seed(100 + s)
        x = torch.poisson(torch.full((b, 2, 2, h, w), 0.3), generator=g)
        st_ref = list(fwd(sd, x, *st_ref, init))
        t0 = time.time()
        if use_step:
            o = m.step(x.cuda(), reset=init)
            torch.cuda.synchronize()
            errs = [(o.cpu() - st_ref[-1]).abs().max().item()]
        else:
            st = list(m(x.cuda(), *st, init))
            torch.cuda.synchronize()
            errs = [(a.cpu() - r).abs().max().item() for a, r in zip(st, st_ref)]
        dt = time.time() - t0
        init = False
        print('  step %d: max|err| %s   ref max: out %.3f hid %.3f   (%.1f ms, %d launches/step)'
              % (s, ' '.join('%.3e' % e for e in errs), st_ref[-1].abs().max().item(), st_ref[0].abs().max().item(),
                 dt * 1e3, m._engine.launches_per_step))


def enc_case():
    import numpy as np
    import torch
    from oracle import encodings_np as E
    from oracle.make_golden import synth_events
    from bmcnet_esr_b200.dataloader import encodings as G
    for (n, h, w, B) in [(2048, 45, 80, 5), (100000, 45, 80, 5), (50000, 180, 320, 3), (30000, 360, 640, 2)]:
        ev = synth_events(n, h, w, 7, oor=0.03, dup=True, frac=True)
        fns = [('channels', lambda a, M: M.events_to_channels(a[0], a[1], a[3], sensor_size=(h, w))),
               ('image', lambda a, M: M.events_to_image(a[0], a[1], a[3], sensor_size=(h, w))),
               ('voxel', lambda a, M: M.events_to_voxel(a[0], a[1], a[2], a[3], B, sensor_size=(h, w))),
               ('stack_pol', lambda a, M: M.events_to_stack_polarity(a[0], a[1], a[2], a[3], B, sensor_size=(h, w))),
               ('stack_nopol', lambda a, M: M.events_to_stack_no_polarity(a[0], a[1], a[2], a[3], B, sensor_size=(h, w))),
               ('voxel_torch', lambda a, M: M.events_to_voxel_torch(a[0], a[1], a[2], a[3], B, sensor_size=(h, w))),
               ('image_bilinear', lambda a, M: M.events_to_image_torch(a[0], a[1], a[3], sensor_size=(h, w), interpolation='bilinear'))]
        for name, fn in fns:
            ca = [x.copy() for x in ev]
            ga = [torch.from_numpy(x.copy()).cuda() for x in ev]
            ref = fn(ca, E)
            got = fn(ga, G)
            torch.cuda.synchronize()
            got = got.cpu().numpy()
            mut = all(np.array_equal(c, g.cpu().numpy()) for c, g in zip(ca, ga))
            print('  n=%d %dx%d %-14s exact=%s max|err| %.3e (ref max %.2f) mutation_match=%s'
                  % (n, h, w, name, np.array_equal(ref, got), np.abs(ref - got).max(), np.abs(ref).max(), mut))


def perf_case():
    import torch
    from bmcnet_esr_b200 import kernels as K
    from bmcnet_esr_b200.dataloader import encodings as G

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    for n in (10 ** 7, 10 ** 8, 4 * 10 ** 8):
        for (h, w) in ((45, 80), (180, 320), (360, 640)):
            xs = torch.rand(n, device='cuda') * w
            ys = torch.rand(n, device='cuda') * h
            ps = torch.randint(0, 2, (n,), device='cuda').float() * 2 - 1
            ms = timeit(lambda: G.events_to_channels(xs, ys, ps, sensor_size=(h, w)), iters=5, warm=2)
            print('  channels n=%.0e %dx%d: %.3f ms  %.1f Gev/s  %.0f GB/s' % (n, h, w, ms, n / ms / 1e6, 12 * n / ms / 1e6))
        ts = torch.sort(torch.rand(n, device='cuda'))[0]
        h, w = 45, 80
        ms = timeit(lambda: G.events_to_voxel(xs * 0 + 1, ys * 0 + 1, ts, ps, 5, sensor_size=(h, w)), iters=3, warm=1)
        del ts
        torch.cuda.empty_cache()
    for b in (1, 4, 16):
        h, w = 45, 80
        x = torch.randn(b, 128, h, w, device='cuda')
        wt = torch.randn(128, 128, 3, 3, device='cuda') * 0.03
        wpk = K.pack_conv_weight(wt, [(0, 128)])
        src = K.pack_nchw(x)
        bias = torch.zeros(128, device='cuda')
        for impl in (0,):
            ms = timeit(lambda: K.conv_gemm([src], wpk, bias, b, h, w, 9, relu=True, impl=impl), iters=20)
            rows = b * K.rows_per_image(h, w)
            print('  conv3x3 128->128 B=%d rows=%d: %.1f us  %.1f TFLOP/s (padded rows) ' % (b, rows, ms * 1e3, 2 * rows * 1152 * 128 / ms / 1e9))


def convperf_case():
    import ctypes as C
    import torch
    from bmcnet_esr_b200 import _lib, kernels as K
    h, w = 45, 80
    for taps, segs in ((9, 1), (1, 1), (1, 2)):
        for jobs in (2,):
            for b in (1, 4, 16, 19):
                for impl in (0, 2):
                    src = [K.pack_nchw(torch.randn(b, 128, h, w, device='cuda')) for _ in range(jobs * segs)]
                    wt = torch.randn(128, 128 * segs, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device='cuda') * 0.03
                    wpk = K.pack_conv_weight(wt, [(i * 128, 128) for i in range(segs)])
                    bias = torch.zeros(128, device='cuda')
                    outs = [torch.empty_like(src[0]) for _ in range(jobs)]
                    jarr = (_lib.GemmJob * jobs)()
                    for j in range(jobs):
                        jarr[j].n_seg = segs
                        for sg in range(segs):
                            t = src[j * segs + sg]
                            jarr[j].a[sg] = t.data_ptr(); jarr[j].a_rows[sg] = t.shape[0]; jarr[j].a_ch[sg] = 128
                        jarr[j].w = wpk.data_ptr(); jarr[j].w_rows = 128; jarr[j].w_k = wpk.shape[0] * 64
                        jarr[j].bias = bias.data_ptr(); jarr[j].out_act16 = outs[j].data_ptr(); jarr[j].relu = 1
                    launch = lambda: _lib.check(_lib.lib().bmc_conv_gemm(jarr, jobs, 128, taps, b, h, w, impl, _lib.stream_ptr()))
                    for _ in range(3):
                        launch()
                    torch.cuda.synchronize()
                    reps = 20
                    g = torch.cuda.CUDAGraph()
                    side = torch.cuda.Stream()
                    with torch.cuda.stream(side):
                        with torch.cuda.graph(g, stream=side):
                            for _ in range(reps):
                                launch()
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) / reps * 1e3
                    fl = 2.0 * 128 * 128 * segs * taps * h * w * b * jobs
                    print('  taps=%d segs=%d jobs=%d B=%2d impl=%d: %7.1f us  %7.1f TFLOP/s (real px)' % (taps, segs, jobs, b, impl, us, fl / us / 1e6))


def run_stage(name):
    import torch
    print('== stage', name, flush=True)
    if name == 'enc':
        enc_case()
    elif name in ('simt_conv', 'tc_conv'):
        impl = 1 if name.startswith('simt') else 0
        conv_case(impl, [128], 128, 1, False, False, False)
        conv_case(impl, [128], 128, 9, False, False, False)
        conv_case(impl, [128], 128, 9, True, True, False)
        conv_case(impl, [128, 128], 128, 1, False, True, False)
        conv_case(impl, [128, 64], 128, 9, True, False, False)
        conv_case(impl, [128], 128, 9, True, True, False, b=3, h=45, w=80)
    elif name == 'tc_conv_ln':
        conv_case(1, [128, 128], 128, 1, False, False, True)
        conv_case(0, [128, 128], 128, 1, False, False, True)
    elif name == 'tc_conv32':
        conv_case(1, [128, 128], 32, 9, False, False, False)
        conv_case(0, [128, 128], 32, 9, False, False, False)
    elif name == 'simt_att':
        att_case(1)
    elif name == 'tc_att':
        att_case(0)
        att_case(0, b=1, h=45, w=80, n_split=8)
    elif name == 'plain_simt':
        model_case(True, True)
    elif name == 'plain_tc':
        model_case(True, False)
    elif name == 'full_simt':
        model_case(False, True, h=12, w=20)
    elif name == 'full_tc':
        model_case(False, False)
    elif name == 'convperf':
        convperf_case()
    elif name == 'perf':
        perf_case()
    elif name == 'step_tc':
        model_case(True, False, use_step=True, steps=4)
        model_case(False, False, use_step=True, steps=4)
    torch.cuda.synchronize()
    print('== stage', name, 'done', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--stage':
        run_stage(sys.argv[2])
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for s in stages:
        env = dict(os.environ)
        name = s
        if '@' in s:                      # stage@VAR=val,VAR2=val2
            name, kv = s.split('@', 1)
            for item in kv.split(','):
                k, v = item.split('=', 1)
                env[k] = v
            print('## env', kv)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--stage', name], timeout=300,
                               capture_output=True, text=True, env=env)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print('!! stage %s exit code %d\n%s' % (s, r.returncode, r.stderr[-3000:]))
        except subprocess.TimeoutExpired:
            print('!! stage %s TIMEOUT' % s)
        sys.stdout.flush()
