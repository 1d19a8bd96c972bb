"""BMCNet -- drop-in for the reference `models/BMCNet.py` (:3-121).

Same constructor, forward signature and state_dict keys (318, with the reference's aliasing:
one ParallelBlk object repeated n_b times, conv2 = conv1, ...); the forward pass is one call
into libbmc_b200 (bmc_model_forward).
"""
import torch
import torch.nn as nn

from . import _train
from ._engine import Engine
from .submodules import BIE, PixelUnShuffle, ResidualBlock_noBN, initialize_weights
from .._lib import MODEL_BMCNET


class ParallelBlk(nn.Module):
    """Parameter container mirroring the reference ParallelBlk (BMCNet.py:3-17)."""

    def __init__(self, nf=64):
        super().__init__()
        self.conv1 = ResidualBlock_noBN(nf)
        self.conv2 = self.conv1
        self.conv1_st = ResidualBlock_noBN(nf)
        self.conv2_st = self.conv1_st
        self.lBIE = BIE(nf)      # local: within one polarity
        self.gBIE = BIE(nf)      # global: across polarities
        initialize_weights([self.conv1, self.conv2, self.conv1_st, self.conv2_st], 0.1)


class Backbone(nn.Module):
    """Parameter container mirroring the reference Backbone (BMCNet.py:35-55)."""

    def __init__(self, n_c, n_b, scale, repeat):
        super().__init__()
        pad = (1, 1)
        self.conv_fpst = nn.Conv2d(scale ** 2 + n_c + 2 * repeat, n_c, 3, 1, padding=pad)
        self.conv_fnst = self.conv_fpst
        self.conv_fps = nn.Conv2d(repeat + n_c, n_c, 3, 1, padding=pad)
        self.conv_fns = self.conv_fps
        self.conv_fs = nn.Conv2d(scale ** 2 * 2 + n_c * 3, n_c, 3, 1, padding=pad)
        self.para_reschunk = nn.ModuleList([ParallelBlk(n_c)] * n_b)
        self.scale = scale
        self.conv_hs = nn.Conv2d(n_c, n_c, 3, 1, padding=pad)
        self.conv_hp = nn.Conv2d(n_c, n_c, 3, 1, padding=pad)
        self.conv_hn = nn.Conv2d(n_c, n_c, 3, 1, padding=pad)
        self.conv_o = nn.Conv2d(n_c * 2, scale ** 2 * 2, 3, 1, padding=pad)
        initialize_weights([self.conv_fpst, self.conv_fnst, self.conv_fps, self.conv_fns, self.conv_fs,
                            self.conv_hs, self.conv_hp, self.conv_hn, self.conv_o], 0.1)


class BMCNet(nn.Module):
    def __init__(self, scale, n_c, n_b, repeat=3):
        super().__init__()
        self.neuro = Backbone(n_c, n_b, scale, repeat=repeat)
        self.scale = scale
        self.down = PixelUnShuffle(scale)
        self.repeat = repeat
        self._engine = Engine(MODEL_BMCNET, scale, n_c, n_b, repeat)
        self.loss_scale = _train.DEFAULT_LOSS_SCALE      # static fp16 loss scale of the training path (models/_train.py)

    def forward(self, x, x_h, x_h_p, x_h_n, x_o, init):
        """Same arguments and return order as the reference (BMCNet.py:95-121):
        (x_h, x_h_p, x_h_n, x_o), x_o = [B,2,sH,sW].  The reference hands (x_h, x_h_p, x_h_n)
        positionally to Backbone.forward(xs, hp, hn, hs, o); that pairing is reproduced."""
        if _train.route(self, x, x_h, x_h_p, x_h_n, x_o):
            # train() mode with autograd recording (train.py:192,202-237): kernels behind autograd Functions
            return _train.forward_full(self, _train.context(self), x, x_h, x_h_p, x_h_n, x_o, init)
        with torch.no_grad():
            (h, hp, hn), o = self._engine.forward(self, x, [x_h, x_h_p, x_h_n], x_o, init)
        return h, hp, hn, o

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
