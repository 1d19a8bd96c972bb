"""Generate `tests/golden/*.npz` from the REFERENCE ITSELF -- TEST INFRASTRUCTURE.

Run in the build container only (it imports `/root/reference`, which does not
exist on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/, oracle/_ref/

The reference has no tests or golden vectors of its own, so
parity is pinned on what its code computes here (torch fp32 CPU, version recorded
in each file).  Inputs are seeded; inputs AND outputs are stored so the fixtures
stay valid if a generator changes.  Also stages the one shipped checkpoint
(`pretrain/BMCNet_plain_nfs_x4.pth`, a weight blob, not source) into the
git-ignored `oracle/_ref/`, which travels to the GPU box with the snapshot.
"""
import hashlib
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('BMC_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)


def synth_events(n, h, w, seed, oor=0.0, dup=False, frac=False):
    """Seeded synthetic events in the reference input contract (base_dataset.py:24-31)."""
    g = np.random.default_rng(seed)
    xs = g.integers(0, w, n).astype(np.float32)
    ys = g.integers(0, h, n).astype(np.float32)
    if frac:
        xs += (g.random(n) * 0.99).astype(np.float32)
        ys += (g.random(n) * 0.99).astype(np.float32)
    if oor > 0:
        k = g.random(n) < oor
        xs[k] = g.choice([-3.0, -0.5, float(w), w + 5.5], int(k.sum())).astype(np.float32)
        k = g.random(n) < oor
        ys[k] = g.choice([-1.0, float(h), h + 2.25], int(k.sum())).astype(np.float32)
    t = np.sort(g.random(n))
    if dup:
        t = np.round(t * 50) / 50
    if n > 0:
        t = (t - t[0]) / (t[-1] - t[0] + 1e-6)
    ts = t.astype(np.float32)
    ps = g.choice([-1.0, 1.0], n).astype(np.float32)
    return xs, ys, ts, ps


ENCODER_CASES = [
    # name, n, (H, W), seed, oor, dup, frac, B
    ('nfs_window', 2048, (45, 80), 11, 0.0, False, False, 5),
    ('eventzoom_window', 1024, (31, 56), 12, 0.0, False, False, 5),
    ('oor_quirk', 3000, (12, 16), 13, 0.08, False, True, 3),
    ('dup_timestamps', 4000, (12, 16), 14, 0.05, True, False, 5),
    ('tiny_n4', 4, (7, 5), 15, 0.0, False, False, 3),
    ('early_out_n3', 3, (7, 5), 16, 0.0, False, False, 3),
    ('single_bin', 500, (9, 11), 17, 0.02, False, True, 1),
    ('dense_20k', 20000, (45, 80), 18, 0.01, True, True, 5),
]


def make_encoder_goldens():
    sys.path.insert(0, REF)
    from dataloader import encodings as R
    for name, n, (h, w), seed, oor, dup, frac, nb in ENCODER_CASES:
        ev = synth_events(n, h, w, seed, oor, dup, frac)
        out = {'n': n, 'h': h, 'w': w, 'B': nb, 'torch': torch.__version__}
        for i, k in enumerate(('xs', 'ys', 'ts', 'ps')):
            out['in_' + k] = ev[i]
        fns = {
            'image': lambda a: R.events_to_image(a[0], a[1], a[3], sensor_size=(h, w)),
            'channels': lambda a: R.events_to_channels(a[0], a[1], a[3], sensor_size=(h, w)),
            'voxel': lambda a: R.events_to_voxel(a[0], a[1], a[2], a[3], nb, sensor_size=(h, w)),
            'image_torch': lambda a: R.events_to_image_torch(a[0], a[1], a[3], sensor_size=(h, w)),
            'image_torch_bilinear': lambda a: R.events_to_image_torch(
                a[0], a[1], a[3], sensor_size=(h, w), interpolation='bilinear'),
            'stack_polarity': lambda a: R.events_to_stack_polarity(
                a[0], a[1], a[2], a[3], nb, sensor_size=(h, w)),
            'stack_no_polarity': lambda a: R.events_to_stack_no_polarity(
                a[0], a[1], a[2], a[3], nb, sensor_size=(h, w)),
            'voxel_torch': lambda a: R.events_to_voxel_torch(
                a[0], a[1], a[2], a[3], nb, sensor_size=(h, w)),
            'voxel_torch_hard': lambda a: R.events_to_voxel_torch(
                a[0], a[1], a[2], a[3], nb, sensor_size=(h, w), temporal_bilinear=False),
        }
        for fname, fn in fns.items():
            args = [torch.from_numpy(a.copy()) for a in ev]
            out['out_' + fname] = fn(args).numpy()
            for i, k in enumerate(('xs', 'ys', 'ts', 'ps')):      # in-place side effects
                if not np.array_equal(args[i].numpy(), ev[i]):    # absent key == unchanged
                    out['mut_%s_%s' % (fname, k)] = args[i].numpy()
        np.savez_compressed(os.path.join(GOLD, 'enc_%s.npz' % name), **out)
        print('enc', name, {k: v.shape for k, v in out.items() if k.startswith('out_')})


def synth_counts(b, h, w, seed, rate=0.3):
    """Seeded [b,2,2,h,w] count frames: the `inp_cnt.transpose(1,2)` of infer_BMCNet.py:50."""
    g = torch.Generator().manual_seed(seed)
    return torch.poisson(torch.full((b, 2, 2, h, w), rate), generator=g)


def sha256(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def make_model_goldens():
    sys.path.insert(0, REF)
    from models.BMCNet import BMCNet
    from models.BMCNet_plain import BMCNet_plain
    from oracle import bmcnet_fp32 as O

    src = os.path.join(REF, 'pretrain', 'BMCNet_plain_nfs_x4.pth')
    dst_dir = os.path.join(ROOT, 'oracle', '_ref')
    os.makedirs(dst_dir, exist_ok=True)
    dst = os.path.join(dst_dir, 'BMCNet_plain_nfs_x4.pth')
    shutil.copyfile(src, dst)
    plain_sd = torch.load(dst, map_location='cpu')

    def rollout(model, n_state, b, h, w, steps, seed):
        st = [torch.zeros(b, 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w)]
        xs, outs = [], []
        init = True
        for s in range(steps):
            x = synth_counts(b, h, w, seed + s)
            with torch.no_grad():
                st = list(model(x, *st, init))
            init = False
            xs.append(x.numpy())
            outs.append(st[-1].numpy())
        return np.stack(xs), np.stack(outs), [t.numpy() for t in st[:-1]]

    # (1) BMCNet_plain with the shipped checkpoint (BASELINE config 2)
    m = BMCNet_plain(4, 128, 5).eval()
    m.load_state_dict(plain_sd, strict=True)
    xs, outs, hid = rollout(m, 2, 1, 16, 24, 3, 100)
    np.savez_compressed(os.path.join(GOLD, 'model_plain_shipped.npz'), x=xs, x_o=outs, x_h=hid[0],
                        ckpt_sha256=sha256(dst), torch=torch.__version__)
    print('plain shipped', xs.shape, outs.shape)

    # (2) BMCNet_plain, seeded surrogate weights (needs no checkpoint on the box)
    ssd = O.surrogate_state_dict(plain=True, seed=2024)
    m = BMCNet_plain(4, 128, 5).eval()
    m.load_state_dict(ssd, strict=True)
    xs, outs, hid = rollout(m, 2, 2, 10, 16, 3, 200)
    np.savez_compressed(os.path.join(GOLD, 'model_plain_surrogate.npz'), x=xs, x_o=outs, x_h=hid[0],
                        seed=2024, torch=torch.__version__)

    # (3) BMCNet, seeded surrogate (+ transplant of the shipped plain tensors)
    for tag, tr in (('surrogate', None), ('transplant', plain_sd)):
        ssd = O.surrogate_state_dict(plain=False, seed=2024, transplant=tr)
        m = BMCNet(4, 128, 5).eval()
        m.load_state_dict(ssd, strict=True)
        xs, outs, hid = rollout(m, 4, 1, 10, 16, 3, 300)
        np.savez_compressed(os.path.join(GOLD, 'model_bmcnet_%s.npz' % tag), x=xs, x_o=outs,
                            x_h=hid[0], x_h_p=hid[1], x_h_n=hid[2], seed=2024,
                            torch=torch.__version__)
        print('bmcnet', tag, xs.shape, outs.shape)

    # (4) state_dict contract: key order, shapes, alias groups
    for plain in (False, True):
        m = (BMCNet_plain if plain else BMCNet)(4, 128, 5)
        sd = m.state_dict()
        ptr = {}
        groups = [ptr.setdefault(v.data_ptr(), len(ptr)) for v in sd.values()]
        np.savez_compressed(os.path.join(GOLD, 'statedict_%s.npz' % ('plain' if plain else 'bmcnet')),
                            keys=np.array(list(sd.keys())),
                            shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
                            alias_group=np.array(groups),
                            n_unique_params=sum(p.numel() for p in m.parameters()))


def synth_recording(n, h, w, seed, oor=0.0, t_base=1.6e9):
    """Seeded raw recording as stored on disk: int16 xs / ys, float64 ts / ps (event_packagers.py:128-156)."""
    g = np.random.default_rng(seed)
    xs = g.integers(0, w, n).astype(np.int16)
    ys = g.integers(0, h, n).astype(np.int16)
    if oor > 0:
        k = g.random(n) < oor
        xs[k] = g.choice([-3, w, w + 5], int(k.sum())).astype(np.int16)
        k = g.random(n) < oor
        ys[k] = g.choice([-1, h, h + 2], int(k.sum())).astype(np.int16)
    ts = np.sort(g.random(n) * 37.5 + t_base)         # float64 seconds; absolute epochs collapse in the float32 cast (base_dataset.py:28)
    ps = g.choice([-1.0, 1.0], n)
    return xs, ys, ts, ps


def make_format_golden():
    """BaseDataset.event_formatting of the reference on raw windows (base_dataset.py:24-31)."""
    from dataloader.base_dataset import BaseDataset
    out = {'torch_version': torch.__version__}
    for name, n, seed, t_base in (('w2048', 2048, 31, 1.6e9), ('rel2048', 2048, 34, 0.25), ('w5', 5, 32, 3.0), ('w1', 1, 33, 0.0)):
        xs, ys, ts, ps = synth_recording(n, 45, 80, seed, oor=0.05, t_base=t_base)
        ev = np.concatenate((xs[np.newaxis], ys[np.newaxis], ts[np.newaxis], ps[np.newaxis]), axis=0)   # get_events, h5dataset.py:407-414
        res = BaseDataset.event_formatting(ev)
        out[name + '_xs'] = xs; out[name + '_ys'] = ys; out[name + '_ts'] = ts; out[name + '_ps'] = ps
        out[name + '_out'] = res.numpy()
    np.savez_compressed(os.path.join(GOLD, 'fmt_events.npz'), **out)


def synth_stack(b, shape, seed, rate=0.15, vmax=4):
    """Seeded count stack with sparse small integers plus fractional noise (the functions round first)."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randint(-vmax, vmax + 1, (b,) + shape, generator=g).float()
    keep = torch.rand((b,) + shape, generator=g) < rate
    return v * keep + (torch.rand((b,) + shape, generator=g) - 0.5) * 0.6


def make_redistribute_golden():
    """The reference's inverse encoders on small stacks (encodings.py:367-464, 653-671)."""
    from dataloader import encodings as R
    out = {'torch_version': torch.__version__}
    pol = synth_stack(3, (2, 4, 6, 7), 41).abs()                       # per-polarity counts are non-negative
    pol[2] = 0                                                         # an empty entry
    nop = synth_stack(3, (5, 6, 7), 42)
    nop[1] = 0
    nop[1, 0, 0, 0], nop[1, 1, 2, 3] = 2.0, -2.0                       # sums to zero: the reference treats it as empty
    out['pol_in'] = pol.numpy(); out['pol_out'] = R.python_event_redistribute_PolarityStack(pol.clone()).numpy()
    out['nop_in'] = nop.numpy(); out['nop_out'] = R.python_event_redistribute_NoPolarityStack(nop.clone()).numpy()
    zero = torch.zeros(2, 3, 4, 5)
    out['zero_out'] = R.python_event_redistribute_NoPolarityStack(zero).numpy()
    out['s2c_in'] = nop.numpy(); out['s2c_out'] = R.stack2cnt(nop.clone()).numpy()
    np.savez_compressed(os.path.join(GOLD, 'redistribute.npz'), **out)


def _stats(t, k=24):
    """Compact fingerprint of a tensor: sum, abs-sum (float64) and its first k values."""
    f = t.detach().double().flatten()
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], f[:k].numpy(), np.zeros(max(0, k - f.numel()))])


def make_train_goldens():
    """Two training iterations of the REFERENCE (its nn.Modules, nn.MSELoss, torch.optim.Adam(amsgrad)) on seeded
    data, the way train.py:202-237 runs them; fingerprints of every unique parameter's gradient (iteration 1) and
    value (after iteration 2) pin oracle/train_step.py."""
    sys.path.insert(0, REF)
    from models.BMCNet import BMCNet
    from models.BMCNet_plain import BMCNet_plain
    from oracle import bmcnet_fp32 as O

    for plain, tag, gt_hw in ((True, 'plain', (40, 64)), (False, 'bmcnet', (38, 62))):     # (38,62): bicubic resize path
        b, h, w, T = 1, 10, 16, 3
        sd = O.surrogate_state_dict(plain=plain, seed=2024)
        m = (BMCNet_plain if plain else BMCNet)(4, 128, 5)
        m.load_state_dict(sd, strict=True)
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-5, amsgrad=True)    # train_nfs.yml:28-34
        mse = torch.nn.MSELoss()
        xs = [synth_counts(b, h, w, 700 + s) for s in range(T)]
        g = torch.Generator().manual_seed(77)
        gts = [torch.poisson(torch.full((b, 2) + gt_hw, 0.3), generator=g) for _ in range(T)]
        names = {id(p): O._alias_root(n) for n, p in m.named_parameters()}
        out = {'x': np.stack([x.numpy() for x in xs]), 'gt': np.stack([t.numpy() for t in gts]), 'losses': []}
        for it in range(2):
            opt.zero_grad()
            loss, init, st = 0, True, None
            for x, gt in zip(xs, gts):                                              # train.py:206-234
                if init:
                    z = torch.zeros_like(x[:, 0:1, 0])
                    st = [z.repeat(1, 128, 1, 1)] * (1 if plain else 3) + [z.repeat(1, 32, 1, 1)]
                st = list(m(x, *st, init))
                init = False
                pred = st[-1]
                if pred.shape[-2:] != gt.shape[-2:]:
                    pred = torch.nn.functional.interpolate(pred, size=gt.shape[-2:], mode='bicubic', align_corners=False)
                loss = loss + mse(pred, gt)
            loss.backward()
            if it == 0:
                for p in m.parameters():
                    out['grad.' + names[id(p)]] = _stats(p.grad)
            opt.step()
            out['losses'].append(loss.item())
        for p in m.parameters():
            out['param.' + names[id(p)]] = _stats(p)
        out['losses'] = np.array(out['losses'])
        np.savez_compressed(os.path.join(GOLD, 'train_step_%s.npz' % tag), torch=torch.__version__, **out)
        print('train golden', tag, out['losses'], len([k for k in out if k.startswith('grad.')]), 'unique parameters')


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    if 'train' in sys.argv[1:]:
        make_train_goldens()
        sys.exit(0)
    make_encoder_goldens()
    make_model_goldens()
    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print('golden bytes', total)
    make_format_golden()
    make_redistribute_golden()
    make_train_goldens()
