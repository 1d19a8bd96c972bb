"""CPU restatement of the reference's training iteration -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/train.py:202-237 (one iteration: zero_grad, an `inputs_seq` of recurrent steps from zero
state, per-step nn.MSELoss against `gt_cnt[:, 1]` after an optional bicubic resize, the SUM of the step losses,
one backward through the whole sequence, one optimiser step) with the optimiser of config/train_nfs.yml:28-34:
Adam(lr=1e-4, weight_decay=1e-5, amsgrad=True), restated from its published update rule instead of calling
torch.optim.  The forward is oracle/bmcnet_fp32.py with autograd switched back on; aliased state_dict keys
 are bound to ONE leaf tensor each, so their gradients accumulate the way the reference's shared
modules do.

Pinned by oracle/make_golden.py::make_train_goldens against the reference's own nn.Modules + nn.MSELoss +
torch.optim.Adam run in this container (tests/golden/train_step_*.npz; tests/test_oracle_vs_golden.py).
The CUDA training path (bmcnet_esr_b200/models/_train.py, csrc/train.cu) is checked against this file in
tests/test_train_graph_cpu.py (host logic, CPU) and tests/test_gpu_train*.py (kernels).  Both reference trainers
(/root/reference/train.py and train_plain.py) run the same iteration.
"""
import torch
import torch.nn.functional as F

from . import bmcnet_fp32 as O

_FWD = {False: O.bmcnet_forward.__wrapped__, True: O.bmcnet_plain_forward.__wrapped__}     # without no_grad


def unique_parameters(sd):
    """{alias root -> leaf tensor (requires_grad)} and the full state_dict view {key -> that leaf}."""
    leaves, view = {}, {}
    for k, v in sd.items():
        root = O._alias_root(k)
        if root not in leaves:
            leaves[root] = sd[root].detach().clone().requires_grad_(True)
        view[k] = leaves[root]
    return leaves, view


def sequence_loss(view, inputs_seq, gts, plain, scale=4, n_c=128):
    """train.py:206-234.  inputs_seq: list of [B,2(T),2(pol),H,W] count stacks (already `transpose(1,2)`-ed, i.e.
    what the model receives); gts: list of [B,2,kH,kW].  Returns (sum of per-step MSE, last step's MSE)."""
    fwd = _FWD[plain]
    loss, mse, state = 0, None, None
    for x, gt in zip(inputs_seq, gts):
        if state is None:
            b, _, _, h, w = x.shape
            zeros = lambda c: torch.zeros(b, c, h, w)
            state = [zeros(n_c)] * (1 if plain else 3) + [zeros(scale * scale * 2)]
            state = list(fwd(view, x, *state, True))
        else:
            state = list(fwd(view, x, *state, False))
        pred = state[-1]
        if pred.shape[-2:] != gt.shape[-2:]:
            pred = F.interpolate(pred, size=gt.shape[-2:], mode='bicubic', align_corners=False)
        mse = F.mse_loss(pred, gt)
        loss = loss + mse
    return loss, mse


def loss_and_grads(sd, inputs_seq, gts, plain):
    """One backward through the whole sequence (train.py:236).  Returns (loss, {root -> grad}, leaves)."""
    leaves, view = unique_parameters(sd)
    loss, _ = sequence_loss(view, inputs_seq, gts, plain)
    roots = sorted(leaves)
    grads = torch.autograd.grad(loss, [leaves[r] for r in roots], allow_unused=True)
    return loss.detach(), {r: (g if g is not None else torch.zeros_like(leaves[r])) for r, g in zip(roots, grads)}, leaves


def adam_amsgrad_step(params, grads, state, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
    """torch.optim.Adam(amsgrad=True) with L2 weight decay, one step, in place on `params` (dicts keyed alike).
    state: {} on the first call; holds step, exp_avg, exp_avg_sq, max_exp_avg_sq per key afterwards.

        g  = grad + wd * p
        m  = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;  vmax = max(vmax, v)
        p -= lr / (1 - b1^t) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
    """
    b1, b2 = betas
    for k, p in params.items():
        st = state.setdefault(k, {'step': 0, 'm': torch.zeros_like(p), 'v': torch.zeros_like(p),
                                  'vmax': torch.zeros_like(p)})
        st['step'] += 1
        t = st['step']
        g = grads[k] + weight_decay * p
        st['m'].mul_(b1).add_(g, alpha=1 - b1)
        st['v'].mul_(b2).addcmul_(g, g, value=1 - b2)
        torch.maximum(st['vmax'], st['v'], out=st['vmax'])
        denom = (st['vmax'].sqrt() / (1 - b2 ** t) ** 0.5).add_(eps)
        p.addcdiv_(st['m'], denom, value=-lr / (1 - b1 ** t))
    return state


def train_iteration(sd, inputs_seq, gts, plain, opt_state=None, **adam):
    """zero_grad + forward sequence + backward + optimiser step.  Returns (loss, grads, new unique params, state)."""
    loss, grads, leaves = loss_and_grads(sd, inputs_seq, gts, plain)
    params = {k: v.detach().clone() for k, v in leaves.items()}
    state = adam_amsgrad_step(params, grads, {} if opt_state is None else opt_state, **adam)
    return loss, grads, params, state
