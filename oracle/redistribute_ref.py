"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product) of the reference's inverse encoders:
python_event_redistribute_PolarityStack / _NoPolarityStack (dataloader/encodings.py:367-464, mode='linear') and
stack2cnt (:653-671).  Third-party arithmetic: torch.round, torch.linspace (float32), Python's stable `sorted`.
Pinned to the reference functions themselves by tests/golden/redistribute.npz (oracle/make_golden.py)."""
import numpy as np
import torch


def _entry_events(entry, num_bins, c_axis):
    """entry: rounded [P,C,Y,X] or [C,Y,X] tensor -> [N,4] events sorted stably by t (encodings.py:384-401 / 433-450)."""
    elist = []
    for ecoor in torch.nonzero(entry):                                   # row-major nonzero order
        value = entry[tuple(ecoor.tolist())]
        num_event = int(torch.abs(value).item())
        el = torch.zeros([num_event, 4])
        el[:, 0] = float(ecoor[-1])
        el[:, 1] = float(ecoor[-2])
        t0 = ecoor[c_axis] / num_bins + 1 / (100 * num_bins)             # :387 / :436
        t1 = (ecoor[c_axis] + 1) / num_bins
        el[:, 2] = torch.linspace(t0, t1, num_event)
        el[:, 3] = 1 if value > 0 else -1
        elist.append(el)
    elist = torch.cat(elist, dim=0)
    order = sorted(range(elist.shape[0]), key=lambda i: elist[i, 2].item())   # Python sorted is stable (:398 / :447)
    return elist[order]


def event_redistribute(event_stack, polarity):
    """The reference function for mode='linear': [B,(2,)C,Y,X] -> [B, maxlen, 4]."""
    batch = event_stack.size()[0]
    num_bins = event_stack.size()[2 if polarity else 1]
    event_stack = event_stack.round()
    out = torch.zeros([batch, 1, 4])
    if event_stack.sum() != 0:
        clouds = []
        for entry in event_stack:
            clouds.append(_entry_events(entry, num_bins, 1 if polarity else 0) if entry.sum() != 0 else torch.zeros([1, 4]))
        maxlen = max(c.size(0) for c in clouds)
        out = torch.zeros((batch, maxlen, 4))
        for b, c in enumerate(clouds):
            out[b, :c.size(0), :] = c
    return out


def stack2cnt(stack):
    stack = stack.clone().detach().round()
    pos, neg = stack.clone(), stack.clone()
    pos[pos < 0] = 0
    neg[neg > 0] = 0
    neg *= -1
    return torch.stack([pos.sum(1), neg.sum(1)], dim=1)


def canonical(cloud, totals):
    """Order-insensitive view of an event cloud for comparisons across timestamp ulps: per entry the multiset of
    (x, y, p, bin-local rank class) is what must agree; returns rows sorted by (t rounded to 1e-5, x, y, p)."""
    out = []
    for b in range(cloud.shape[0]):
        e = np.asarray(cloud[b][:totals[b]], dtype=np.float64)
        key = np.lexsort((e[:, 3], e[:, 1], e[:, 0], np.round(e[:, 2] * 1e5)))
        out.append(e[key])
    return out
