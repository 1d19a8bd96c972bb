"""Time the 3x3 128->128 slab convolution (product kernel) with CUDA events, launches back to back
in a graph.   python tools/time_conv.py B jobs [reps]      (BMC_SLABT_GRID=n caps the grid)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bmcnet_esr_b200 import _lib, kernels as K

b, jobs = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
h, w = 45, 80
src = K.pack_nchw(torch.randn(jobs * b, 128, h, w, device='cuda'))
rows_job = src.shape[0] // jobs
wpk = K.pack_conv_weight(torch.randn(128, 128, 3, 3, device='cuda') * 0.03, [(0, 128)])
bias = torch.zeros(128, device='cuda')
outs = torch.empty_like(src)
jarr = (_lib.GemmJob * jobs)()
for j in range(jobs):
    jarr[j].n_seg = 1
    jarr[j].a[0] = src.data_ptr(); jarr[j].a_rows[0] = src.shape[0]; jarr[j].a_ch[0] = 128
    jarr[j].a_row_base[0] = j * rows_job
    jarr[j].w = wpk.data_ptr(); jarr[j].w_rows = 128; jarr[j].w_k = 1152
    jarr[j].bias = bias.data_ptr(); jarr[j].out_act16 = outs.data_ptr(); jarr[j].out_row_base = j * rows_job
    jarr[j].relu = 1
launch = lambda: _lib.check(_lib.lib().bmc_conv_gemm(jarr, jobs, 128, 9, b, h, w, 0, _lib.stream_ptr()))
for _ in range(3):
    launch()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            launch()
g.replay(); torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); e.record(); torch.cuda.synchronize()
us = a.elapsed_time(e) / reps * 1e3
tiles = jobs * ((src.shape[0] // jobs + 255) // 256)
grid = int(os.environ.get('BMC_SLABT_GRID', '0')) or 148
grid = min(grid, tiles, 148)
rounds = -(-tiles // grid)
tf = 2.0 * 147456 * h * w * b * jobs / (us * 1e-6) / 1e12
print('conv3x3 B=%d jobs=%d grid=%d tiles=%d rounds=%d : %.1f us/launch, %.2f us/tile-round, %.0f TFLOP/s' % (b, jobs, grid, tiles, rounds, us, us / rounds, tf))
