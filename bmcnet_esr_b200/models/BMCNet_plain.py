"""BMCNet_plain -- drop-in for the reference `models/BMCNet_plain.py` (:3-68).

Same constructor, forward signature and state_dict keys (120, with the reference's aliasing);
the forward pass is one call into libbmc_b200 (bmc_model_forward): NHWC fp16 (default build) tcgen05 kernels
with fp32 accumulation, one CUDA graph per step.
"""
import torch
import torch.nn as nn

from . import _train
from ._engine import Engine
from .submodules import BIE, PixelUnShuffle, initialize_weights
from .._lib import MODEL_BMCNET_PLAIN


class Backbone(nn.Module):
    """Parameter container mirroring the reference Backbone (BMCNet_plain.py:3-22)."""

    def __init__(self, n_c, n_b, scale, repeat):
        super().__init__()
        self.conv_f1 = nn.Conv2d(scale ** 2 + n_c + 2 * repeat, n_c, 3, 1, padding=(1, 1))
        self.conv_f2 = self.conv_f1
        self.conv_fs = nn.Conv2d(scale ** 2 * 2 + n_c + 2 * 2 * repeat, n_c, 3, 1, padding=(1, 1))
        self.para_reschunk = nn.ModuleList([BIE(n_c)] * n_b)      # ONE block applied n_b times
        self.scale = scale
        self.conv_h = nn.Conv2d(n_c, n_c, 3, 1, padding=(1, 1))
        self.conv_o = nn.Conv2d(n_c * 2, scale ** 2 * 2, 3, 1, padding=(1, 1))
        # conv_fs keeps PyTorch's default init, as in the reference (BMCNet_plain.py:17)
        initialize_weights([self.conv_f1, self.conv_f2, self.conv_h, self.conv_o], 0.1)


class BMCNet_plain(nn.Module):
    def __init__(self, scale, n_c, n_b, repeat=3):
        super().__init__()
        self.neuro = Backbone(n_c, n_b, scale, repeat=repeat)
        self.scale = scale
        self.down = PixelUnShuffle(scale)
        self.repeat = repeat
        self._engine = Engine(MODEL_BMCNET_PLAIN, scale, n_c, n_b, repeat)
        self.loss_scale = _train.DEFAULT_LOSS_SCALE      # static fp16 loss scale of the training path (models/_train.py)

    def forward(self, x, x_h, x_o, init):
        """x [B,2,T,H,W] counts (frames 0,1 used); x_h [B,n_c,H,W]; x_o [B,2*scale^2,H,W] when
        `init` else the previous [B,2,sH,sW] output.  Returns (x_h, x_o) like the reference."""
        if _train.route(self, x, x_h, x_o):
            # train() mode with autograd recording (train_plain.py:151): kernels behind autograd Functions
            return _train.forward_plain(self, _train.context(self), x, x_h, x_o, init)
        with torch.no_grad():
            (h,), o = self._engine.forward(self, x, [x_h], x_o, init)
        return h, o

    # device-resident recurrence for throughput (not part of the reference API)
    # -- weights / state tracking (models/_engine.py): edits the engine cannot see by itself
    def refresh_weights(self):
        """Re-pack the weights (and forget the resident recurrent state) on the next call.  Needed only after writes
        that bypass the Parameters' version counters, i.e. through `param.data` (`p.data.mul_()`, `p.data.copy_()`,
        EMA / clamp code); `load_state_dict`, `.to()`, optimiser steps and in-place ops on the Parameters are
        detected automatically."""
        self._engine.invalidate()

    @property
    def resident_state_fast_path(self):
        return self._engine.fast_path

    @resident_state_fast_path.setter
    def resident_state_fast_path(self, on):
        self._engine.fast_path = bool(on)
        self._engine._last = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._engine.invalidate()
        return out

    def step(self, x, reset=False, want_output=True):
        with torch.no_grad():
            return self._engine.step(self, x, reset, want_output)
