"""Launch one conv-gemm configuration a few times (for ncu).  python tools/prof_conv.py impl B jobs taps [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bmcnet_esr_b200 import _lib, kernels as K

impl, b, jobs, taps = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
h, w = 45, 80
k = 3 if taps == 9 else 1
src = [K.pack_nchw(torch.randn(b, 128, h, w, device='cuda')) for _ in range(jobs)]
wpk = K.pack_conv_weight(torch.randn(128, 128, k, k, device='cuda') * 0.03, [(0, 128)])
bias = torch.zeros(128, device='cuda')
outs = [torch.empty_like(s) for s in src]
jarr = (_lib.GemmJob * jobs)()
for j in range(jobs):
    jarr[j].n_seg = 1
    jarr[j].a[0] = src[j].data_ptr(); jarr[j].a_rows[0] = src[j].shape[0]; jarr[j].a_ch[0] = 128
    jarr[j].w = wpk.data_ptr(); jarr[j].w_rows = 128; jarr[j].w_k = wpk.shape[0] * 64
    jarr[j].bias = bias.data_ptr(); jarr[j].out_act16 = outs[j].data_ptr(); jarr[j].relu = 1
for _ in range(reps):
    _lib.check(_lib.lib().bmc_conv_gemm(jarr, jobs, 128, taps, b, h, w, impl, _lib.stream_ptr()))
torch.cuda.synchronize()
print('done')
