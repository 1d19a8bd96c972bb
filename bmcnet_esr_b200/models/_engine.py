"""Host glue between the nn.Module mirrors and the C-ABI model object (bmc_model_*).

One Engine per module instance.  It owns (as torch tensors, so the caching allocator stays in
charge) the repacked 16-bit weight buffer and the activation arena, and re-creates them when the
parameters, the device or the problem size change:
  * parameters are re-packed whenever any Parameter's (data_ptr, _version) changed -- i.e. after
    load_state_dict, .to(device), an optimiser step or any in-place op on the Parameter itself -- and whenever
    the owning module calls `invalidate()` (it does after `load_state_dict` and every `_apply`: `.to`, `.cuda`,
    `.float`, ...).
    NOT detected: writes through `param.data` (`p.data.mul_(2)`, `p.data.copy_(...)`, EMA / clamp code): `.data`
    does not share the Parameter's version counter.  After such an edit call `model.refresh_weights()`.
  * the arena / plan are rebuilt when (B, H, W) changes.
The same caveat holds for the resident-state fast path of `forward` (the recurrent state is not re-packed when the
call is handed back the previous call's own, untouched output tensors): an edit of those tensors through `.data` is
not seen; set `model.resident_state_fast_path = False` (or call `model.refresh_weights()`, which also forgets the
resident state) if a caller does that.
"""
import ctypes as C

import torch

from .. import _lib
from .._lib import check, lib, stream_ptr


class Engine:
    def __init__(self, kind, scale, n_c, n_b, repeat):
        self.args = (kind, scale, n_c, n_b, repeat)
        self.handle = None
        self.device = None
        self.weight_buf = None
        self.workspace = None
        self.shape = None
        self.param_sig = None
        self.debug_simt = False
        self._last = None          # the tensors returned by the previous forward() (and their versions)
        self.fast_path = True      # resident-state fast path of forward() (see the module docstring)

    def invalidate(self):
        """Forget the packed weights and the resident recurrent state: the next call re-packs both."""
        self.param_sig = None
        self._last = None

    # -- lifetime -------------------------------------------------------------------------
    def _ensure_handle(self):
        if self.handle is None:
            h = lib().bmc_model_create(*self.args)
            if not h:
                raise _lib.BmcError(lib().bmc_last_error().decode())
            self.handle = C.c_void_p(h)

    def close(self):
        if self.handle is not None:
            lib().bmc_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # a copied / unpickled module gets a fresh engine (the C handle is not shareable)
    def __deepcopy__(self, memo):
        return Engine(*self.args)

    def __reduce__(self):
        return (Engine, self.args)

    # -- weights --------------------------------------------------------------------------
    def _sync_weights(self, module):
        # cheap per-call change detection over the unique Parameters (54 / 24 of them):
        # data_ptr catches .to(device) / load into new storage, _version catches in-place updates
        sig = tuple((p.data_ptr(), p._version) for p in module.parameters())
        if sig == self.param_sig:
            return
        sd = module.state_dict()
        dev = next(iter(sd.values())).device
        if dev.type != 'cuda':
            raise _lib.BmcError('model parameters are on %s: move the module to a CUDA device '
                                '(there is no CPU fallback)' % dev)
        self._ensure_handle()
        if self.device != dev:
            self.workspace, self.shape = None, None
            self.device = dev
        names = list(sd.keys())
        tens = [sd[k].detach().contiguous().float() for k in names]
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tens])
        c_numel = (C.c_int64 * n)(*[t.numel() for t in tens])
        nbytes = lib().bmc_model_weight_bytes(self.handle)
        # fresh buffer each time: a graph replay of an earlier step may still read the old one
        buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        ptr = (buf.data_ptr() + 1023) // 1024 * 1024
        with torch.cuda.device(dev):
            check(lib().bmc_model_load_state_dict(self.handle, c_names, c_ptrs, c_numel, n, C.c_void_p(ptr),
                                                  nbytes, stream_ptr()))
        self.weight_buf = buf
        self.param_sig = sig

    # -- geometry -------------------------------------------------------------------------
    def _sync_shape(self, b, h, w):
        if self.shape == (b, h, w) and self.workspace is not None:
            return
        with torch.cuda.device(self.device):
            check(lib().bmc_model_configure(self.handle, b, h, w))
            nbytes = lib().bmc_model_workspace_bytes(self.handle)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            torch.cuda.current_stream().synchronize()
            check(lib().bmc_model_bind_workspace(self.handle, C.c_void_p(ptr), nbytes))
            check(lib().bmc_model_set_debug_simt(self.handle, int(self.debug_simt)))
        self.workspace = ws
        self.shape = (b, h, w)

    def prepare(self, module, x):
        if not x.is_cuda:
            raise _lib.BmcError('input is on %s: bmcnet_esr_b200 runs on CUDA tensors only' % x.device)
        if x.dim() != 5 or x.shape[1] < 2 or x.shape[2] < 2:
            raise ValueError('x must be [B, 2, T>=2, H, W], got %s' % (tuple(x.shape),))
        self._sync_weights(module)
        if x.device != self.device:
            raise _lib.BmcError('input on %s but parameters on %s' % (x.device, self.device))
        self._sync_shape(x.shape[0], x.shape[3], x.shape[4])

    @property
    def launches_per_step(self):
        return lib().bmc_model_launches_per_step(self.handle)

    def set_debug_simt(self, enable):
        self.debug_simt = bool(enable)
        if self.handle is not None and self.workspace is not None:
            check(lib().bmc_model_set_debug_simt(self.handle, int(self.debug_simt)))

    # -- one step -------------------------------------------------------------------------
    @staticmethod
    def _strides(x):
        return (C.c_int64 * 5)(*x.stride())

    def forward(self, module, x, hiddens, x_o, init):
        """hiddens: 1 (plain) or 3 (BMCNet) fp32 [B,128,H,W] tensors.  Returns (hiddens', x_o')."""
        x = x if x.dtype == torch.float32 else x.float()
        self.prepare(module, x)
        b, _, _, h, w = x.shape
        x_o_in = x_o
        hs = [t.contiguous().float() for t in hiddens]
        x_o = x_o.contiguous().float()
        want = (b, 32, h, w) if init else (b, 2, 4 * h, 4 * w)
        if tuple(x_o.shape) != want:
            raise ValueError('x_o has shape %s, expected %s for init=%s' % (tuple(x_o.shape), want, bool(init)))
        for t in hs:
            if tuple(t.shape) != (b, 128, h, w):
                raise ValueError('hidden state has shape %s, expected %s' % (tuple(t.shape), (b, 128, h, w)))
        outs = [torch.empty_like(t) for t in hs]
        out_o = torch.empty(b, 2, 4 * h, 4 * w, dtype=torch.float32, device=x.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        # The reference loop feeds every call the previous call's outputs (infer_BMCNet.py:61-64).  When the
        # arguments ARE those tensors, untouched, the states are still resident on the device: skip re-packing.
        given = list(hiddens) + [x_o_in]
        resident = self.fast_path and (not init) and self._last is not None and self._last[0] == (self.shape, self.param_sig) and \
            len(given) == len(self._last[1]) and all(a is b and a._version == v for a, (b, v) in zip(given, self._last[1]))
        hp = [None] * 3 if resident else [p(t) for t in hs] + [None] * (3 - len(hs))
        op = [p(t) for t in outs] + [None] * (3 - len(outs))
        with torch.cuda.device(self.device):
            check(lib().bmc_model_forward(self.handle, p(x), self._strides(x), hp[0], hp[1], hp[2],
                                          None if resident else p(x_o), int(bool(init)), op[0], op[1], op[2], p(out_o),
                                          stream_ptr()))
        # strong references: the memory cannot be recycled for another tensor while we compare by identity
        self._last = ((self.shape, self.param_sig), [(t, t._version) for t in outs + [out_o]])
        return outs, out_o

    def step(self, module, x, reset, want_output=True):
        """Device-resident recurrence (bmc_model_step): states never leave the arena."""
        x = x if x.dtype == torch.float32 else x.float()
        self.prepare(module, x)
        b, _, _, h, w = x.shape
        out_o = torch.empty(b, 2, 4 * h, 4 * w, dtype=torch.float32, device=x.device) if want_output else None
        self._last = None
        with torch.cuda.device(self.device):
            check(lib().bmc_model_step(self.handle, C.c_void_p(x.data_ptr()), self._strides(x), int(bool(reset)),
                                       C.c_void_p(out_o.data_ptr()) if want_output else None, stream_ptr()))
        return out_o
