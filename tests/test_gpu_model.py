"""GPU parity of the full forward pass (bmc_model_forward / bmc_model_step through the drop-in
modules) against the CPU oracle (oracle/bmcnet_fp32.py, itself pinned to the reference) and the
reference-generated goldens.

Bars: SR output max-abs <= 1e-2 and PSNR difference <= 0.05 dB against the fp32
reference forward on the same inputs and weights; because the absolute bar is loose for
small-residual weights, hidden states and the learned residual must also be within
1 % of their own max-abs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bmcnet_fp32 as O
from oracle.make_golden import synth_counts

pytestmark = pytest.mark.gpu

ABS_TOL = 1e-2
PSNR_TOL_DB = 0.05
REL_TOL = 1e-2


def _psnr(a, gt):
    mse = ((a - gt) ** 2).mean().item()
    return 10 * np.log10(gt.max().item() ** 2 / mse)


def _models():
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    return BMCNet, BMCNet_plain


def _check_step(got, ref, x, tag):
    """got / ref: [hiddens..., x_o].  Returns the measured errors (for the error-vs-step curves)."""
    o, ro = got[-1].cpu(), ref[-1]
    err = (o - ro).abs().max().item()
    assert err <= ABS_TOL, (tag, 'x_o max-abs', err)
    up = F.interpolate(x[:, :, 1], scale_factor=4, mode='bilinear', align_corners=False)
    res = (ro - up).abs().max().item()
    # the learned residual must be resolved to 1 % of its own size; weights with a residual below 0.02 would make
    # this bar tighter than fp16 storage of the O(1) prediction allows (2^-11 relative), hence the floor
    assert err <= REL_TOL * max(res, 0.02), (tag, 'residual-relative', err, res)
    gt = torch.poisson(F.interpolate(x[:, :, 1], scale_factor=4, mode='nearest') / 16 + 0.05,
                       generator=torch.Generator().manual_seed(1))
    dpsnr = abs(_psnr(o, gt) - _psnr(ro, gt))
    assert dpsnr <= PSNR_TOL_DB, (tag, 'psnr')
    hid = []
    for i, (h, rh) in enumerate(zip(got[:-1], ref[:-1])):
        e = (h.cpu() - rh).abs().max().item()
        assert e <= REL_TOL * rh.abs().max().item(), (tag, 'hidden %d' % i, e, rh.abs().max().item())
        hid.append(e / rh.abs().max().item())
    return {'x_o_max_abs': err, 'residual_max_abs': res, 'psnr_diff_db': dpsnr, 'hidden_rel': hid}


def _rollout(model, fwd, sd, b, h, w, steps, seed, tag, transposed_input=False, curve=None):
    n_state = 2 if fwd is O.bmcnet_plain_forward else 4
    ref = [torch.zeros(b, 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w)]
    got = [t.cuda() for t in ref]
    init = True
    for s in range(steps):
        x = synth_counts(b, h, w, seed + s)
        ref = list(fwd(sd, x, *ref, init))
        xg = x.cuda()
        if transposed_input:          # the reference caller passes inp_cnt.transpose(1, 2) (infer_BMCNet.py:50)
            xg = x.transpose(1, 2).contiguous().cuda().transpose(1, 2)
            assert not xg.is_contiguous()
        got = list(model(xg, *got, init))
        init = False
        m = _check_step(got, ref, x, '%s step %d' % (tag, s))
        if curve is not None:
            curve.append(dict(m, step=s))
    return got, ref


def _save_curve(name, curve):
    """Error-vs-step curves of the long rollouts: written to gpurun_out/ on the GPU box (merged back by gpurun),
    summarised in profiles/r02_rollout_error.md."""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, 'rollout_error_%s.json' % name), 'w') as f:
            json.dump(curve, f)
    except OSError:
        pass


def test_plain_shipped_checkpoint_nfs_shape(plain_ckpt):
    """BASELINE config 2: BMCNet_plain + pretrain/BMCNet_plain_nfs_x4.pth at the NFS LR size."""
    _, BMCNet_plain = _models()
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    m = m.cuda().eval()
    _rollout(m, O.bmcnet_plain_forward, plain_ckpt, 1, 45, 80, 4, 500, 'plain/shipped/45x80', transposed_input=True)
    _rollout(m, O.bmcnet_plain_forward, plain_ckpt, 3, 31, 56, 2, 600, 'plain/shipped/31x56 B=3')


def test_plain_shipped_checkpoint_vs_reference_golden(golden_dir, plain_ckpt):
    _, BMCNet_plain = _models()
    g = np.load(os.path.join(golden_dir, 'model_plain_shipped.npz'))
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(g['x'])
    h, o = torch.zeros(1, 128, 16, 24).cuda(), torch.zeros(1, 32, 16, 24).cuda()
    for s in range(x.shape[0]):
        h, o = m(x[s].cuda(), h, o, s == 0)
        assert (o.cpu() - torch.from_numpy(g['x_o'][s])).abs().max().item() <= ABS_TOL
    assert (h.cpu() - torch.from_numpy(g['x_h'])).abs().max().item() <= REL_TOL * float(np.abs(g['x_h']).max())


@pytest.mark.parametrize('tag', ['surrogate', 'transplant'])
def test_bmcnet_vs_reference_golden(golden_dir, tag, request):
    BMCNet, _ = _models()
    g = np.load(os.path.join(golden_dir, 'model_bmcnet_%s.npz' % tag))
    tr = request.getfixturevalue('plain_ckpt') if tag == 'transplant' else None
    sd = O.surrogate_state_dict(plain=False, seed=int(g['seed']), transplant=tr)
    m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(g['x'])
    st = [torch.zeros(1, 128, 10, 16).cuda() for _ in range(3)] + [torch.zeros(1, 32, 10, 16).cuda()]
    for s in range(x.shape[0]):
        st = list(m(x[s].cuda(), *st, s == 0))
        assert (st[-1].cpu() - torch.from_numpy(g['x_o'][s])).abs().max().item() <= ABS_TOL
    for t, k in zip(st[:3], ('x_h', 'x_h_p', 'x_h_n')):
        assert (t.cpu() - torch.from_numpy(g[k])).abs().max().item() <= REL_TOL * float(np.abs(g[k]).max()), k


def test_bmcnet_surrogate_full_shapes(plain_ckpt):
    """BASELINE configs 1 / 4 shapes with the surrogate weight set (the BMCNet checkpoints are
    not shipped): trained BIE / head tensors transplanted from the plain checkpoint."""
    BMCNet, _ = _models()
    sd = O.surrogate_state_dict(plain=False, seed=7, transplant=plain_ckpt)
    m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    _rollout(m, O.bmcnet_forward, sd, 1, 45, 80, 3, 700, 'bmcnet/transplant/45x80', transposed_input=True)
    _rollout(m, O.bmcnet_forward, sd, 2, 31, 56, 3, 800, 'bmcnet/transplant/31x56 B=2')


def test_device_resident_step_equals_forward(plain_ckpt):
    """bmc_model_step keeps the recurrent state in the arena; same outputs as forward()."""
    BMCNet, BMCNet_plain = _models()
    for cls, sd, n_state in ((BMCNet_plain, plain_ckpt, 2),
                             (BMCNet, O.surrogate_state_dict(plain=False, seed=3, transplant=plain_ckpt), 4)):
        m = cls(4, 128, 5)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        b, h, w = 2, 22, 40
        st = [torch.zeros(b, 128, h, w).cuda() for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w).cuda()]
        xs = [synth_counts(b, h, w, 900 + s).cuda() for s in range(4)]
        outs = []
        for s, x in enumerate(xs):
            st = list(m(x, *st, s == 0))
            outs.append(st[-1].clone())
        for s, x in enumerate(xs):
            o = m.step(x, reset=(s == 0))
            # forward() hands the fp32 hidden state back through the API and re-rounds it, step()
            # keeps the 16-bit copy: identical values, so identical results
            assert (o - outs[s]).abs().max().item() <= 1e-5, (cls.__name__, s)


def test_weights_follow_parameter_updates(plain_ckpt):
    """The repacked copy must track load_state_dict / in-place parameter edits."""
    _, BMCNet_plain = _models()
    m = BMCNet_plain(4, 128, 5).cuda().eval()
    x = synth_counts(1, 12, 20, 1).cuda()
    h, o = torch.zeros(1, 128, 12, 20).cuda(), torch.zeros(1, 32, 12, 20).cuda()
    a = m(x, h, o, True)[1].clone()
    m.load_state_dict(plain_ckpt, strict=True)
    b = m(x, h, o, True)[1].clone()
    assert (a - b).abs().max().item() > 1e-3
    ref = O.bmcnet_plain_forward(plain_ckpt, x.cpu(), h.cpu(), o.cpu(), True)[1]
    assert (b.cpu() - ref).abs().max().item() <= ABS_TOL
    with torch.no_grad():
        m.neuro.conv_o.bias.add_(0.5)
    c = m(x, h, o, True)[1]
    assert abs((c - b).mean().item() - 0.5) < 1e-3


def test_simt_cross_check_path_agrees(plain_ckpt):
    """The on-device SIMT kernels and the tcgen05 kernels implement the same arithmetic."""
    _, BMCNet_plain = _models()
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    m = m.cuda().eval()
    x = synth_counts(2, 12, 20, 2).cuda()
    h, o = torch.zeros(2, 128, 12, 20).cuda(), torch.zeros(2, 32, 12, 20).cuda()
    a = m(x, h, o, True)
    m._engine.set_debug_simt(True)
    b = m(x, h, o, True)
    m._engine.set_debug_simt(False)
    assert (a[1] - b[1]).abs().max().item() <= 2e-3
    assert (a[0] - b[0]).abs().max().item() <= 2e-3 * a[0].abs().max().item() + 1e-3


def test_resident_state_fast_path_equals_repacking(plain_ckpt):
    """forward() skips re-packing the recurrent state when it is handed back its own, untouched outputs
    (the loop of infer_BMCNet.py:61-64); clones of the same values must give bit-identical results."""
    BMCNet, BMCNet_plain = _models()
    for cls, sd, n_state in ((BMCNet_plain, plain_ckpt, 2),
                             (BMCNet, O.surrogate_state_dict(plain=False, seed=5, transplant=plain_ckpt), 4)):
        m = cls(4, 128, 5)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        b, h, w = 2, 20, 33
        xs = [synth_counts(b, h, w, 300 + s).cuda() for s in range(4)]
        z = [torch.zeros(b, 128, h, w).cuda() for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w).cuda()]
        st, fast = list(z), []
        for s, x in enumerate(xs):
            st = list(m(x, *st, s == 0))                      # own outputs handed straight back
            fast.append([t.clone() for t in st])
        st = list(z)
        for s, x in enumerate(xs):
            st = list(m(x, *[t.clone() for t in st], s == 0))  # same values in new tensors: full re-pack
            for a, r in zip(st, fast[s]):
                assert torch.equal(a, r), (cls.__name__, s)
        # an in-place edit of a returned tensor must be seen (version bump -> re-pack)
        st = list(m(xs[0], *z, True))
        st[0].mul_(0.5)
        a = m(xs[1], *st, False)
        bb = m(xs[1], *[t.clone() for t in st], False)
        assert torch.equal(a[-1], bb[-1]) and not torch.equal(a[-1], fast[1][-1])


def test_sequences_of_a_batch_are_independent_at_bench_size(plain_ckpt):
    """Inference shards by independent sequences: at the bench batch (57 sequences, tiles
    straddling images, per-image attention partials spread over CTAs) every sequence must get what it gets when
    stepped alone -- up to the summation order of the attention partials."""
    _, BMCNet_plain = _models()
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    m = m.cuda().eval()
    b, h, w = 57, 45, 80
    xs = [synth_counts(b, h, w, 100 + s).cuda() for s in range(2)]
    st = [torch.zeros(b, 128, h, w).cuda(), torch.zeros(b, 32, h, w).cuda()]
    init = True
    for x in xs:
        st = list(m(x, *st, init))
        init = False
    for pick in (0, 31, 56):
        s1 = [torch.zeros(1, 128, h, w).cuda(), torch.zeros(1, 32, h, w).cuda()]
        init = True
        for x in xs:
            s1 = list(m(x[pick:pick + 1], *s1, init))
            init = False
        for full, one in zip(st, s1):
            err = (full[pick:pick + 1] - one).abs().max().item()
            assert err <= 2e-3 * max(1.0, one.abs().max().item()), (pick, err)


def test_bmcnet_sequences_independent_many_tiles(plain_ckpt):
    """Full BMCNet at a batch with many tiles per CTA: launches of 4 jobs with two weight sets, so the two tiles a
    CTA of conv_slab2_tc has in flight read different weights at job boundaries (non-shared weight slots)."""
    BMCNet, _ = _models()
    sd = O.surrogate_state_dict(plain=False, transplant=plain_ckpt)
    m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    b, h, w = 38, 45, 80
    xs = [synth_counts(b, h, w, 200 + s).cuda() for s in range(2)]

    def run(sel):
        n = len(sel)
        st = [torch.zeros(n, 128, h, w).cuda() for _ in range(3)] + [torch.zeros(n, 32, h, w).cuda()]
        init = True
        for x in xs:
            st = list(m(x[sel], *st, init))
            init = False
        return st

    full = run(list(range(b)))
    for pick in (0, 20, 37):
        one = run([pick])
        for f, o in zip(full, one):
            err = (f[pick:pick + 1] - o).abs().max().item()
            assert err <= 2e-3 * max(1.0, o.abs().max().item()), (pick, err)
    # and against the fp32 oracle for one of the sequences (2 steps)
    ref = [torch.zeros(1, 128, h, w) for _ in range(3)] + [torch.zeros(1, 32, h, w)]
    init = True
    for x in xs:
        ref = list(O.bmcnet_forward(sd, x[20:21].cpu(), *ref, init))
        init = False
    _check_step([t[20:21] for t in full], ref, xs[-1][20:21].cpu(), 'bmcnet B=38 seq 20')


# ---------------------------------------------------------------------------------------------------------------
# Long rollouts ("S >= 16 for parity"; infer_BMCNet.py:46-68 carries the state over a whole recording):
# the hidden state is stored in fp16 between steps and every step runs 5 (plain) / 15 (BMCNet) softmaxes, which is
# where slow drift would hide.  All bars are asserted at EVERY one of the 32 steps.
LONG_STEPS = 32


def test_long_rollout_plain_shipped_nfs(plain_ckpt):
    _, BMCNet_plain = _models()
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    m = m.cuda().eval()
    curve = []
    _rollout(m, O.bmcnet_plain_forward, plain_ckpt, 1, 45, 80, LONG_STEPS, 5000, 'plain/shipped/45x80/S32',
             transposed_input=True, curve=curve)
    _save_curve('plain_shipped_45x80', curve)


@pytest.mark.parametrize('h,w', [(45, 80), (31, 56)])
def test_long_rollout_bmcnet_transplant(plain_ckpt, h, w):
    BMCNet, _ = _models()
    sd = O.surrogate_state_dict(plain=False, seed=7, transplant=plain_ckpt)
    m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    curve = []
    _rollout(m, O.bmcnet_forward, sd, 1, h, w, LONG_STEPS, 6000 + h, 'bmcnet/transplant/%dx%d/S32' % (h, w),
             transposed_input=True, curve=curve)
    _save_curve('bmcnet_transplant_%dx%d' % (h, w), curve)


@pytest.mark.parametrize('kind,b,h,w', [('plain', 95, 45, 80), ('full', 76, 45, 80), ('full', 156, 31, 56)])
def test_bench_batch_sequences_vs_oracle(plain_ckpt, kind, b, h, w):
    """The batches bench.py runs (plain B=95, BMCNet B=76 / 156): three sequences of the batch -- first, middle,
    last -- against the fp32 oracle over 4 recurrent steps, every bar at every step."""
    BMCNet, BMCNet_plain = _models()
    if kind == 'plain':
        cls, sd, fwd, n_state = BMCNet_plain, plain_ckpt, O.bmcnet_plain_forward, 2
    else:
        cls, fwd, n_state = BMCNet, O.bmcnet_forward, 4
        sd = O.surrogate_state_dict(plain=False, transplant=plain_ckpt)
    m = cls(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    picks = [0, b // 2, b - 1]
    steps = 4
    xs = [synth_counts(b, h, w, 7000 + s) for s in range(steps)]
    st = [torch.zeros(b, 128, h, w).cuda() for _ in range(n_state - 1)] + [torch.zeros(b, 32, h, w).cuda()]
    ref = [torch.zeros(len(picks), 128, h, w) for _ in range(n_state - 1)] + [torch.zeros(len(picks), 32, h, w)]
    for s, x in enumerate(xs):
        st = list(m(x.cuda(), *st, s == 0))
        ref = list(fwd(sd, x[picks], *ref, s == 0))
        _check_step([t[picks] for t in st], ref, x[picks], '%s B=%d step %d' % (kind, b, s))


def test_forward_is_bit_reproducible(plain_ckpt):
    """Two runs of the same rollout give bit-identical outputs: the attention partial sums are reduced in a fixed
    order (att_fold sums the per-CTA slots in slot order), nothing in the model path uses float atomics."""
    BMCNet, _ = _models()
    sd = O.surrogate_state_dict(plain=False, transplant=plain_ckpt)
    m = BMCNet(4, 128, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    b, h, w = 5, 45, 80
    xs = [synth_counts(b, h, w, 8000 + s).cuda() for s in range(3)]

    def run():
        st = [torch.zeros(b, 128, h, w).cuda() for _ in range(3)] + [torch.zeros(b, 32, h, w).cuda()]
        for s, x in enumerate(xs):
            st = list(m(x, *st, s == 0))
        return [t.clone() for t in st]

    a, c = run(), run()
    for t, u in zip(a, c):
        assert torch.equal(t, u)
