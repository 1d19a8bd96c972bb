"""Batch-1 latency (infer_BMCNet.py:54-68 timing) of both models: python tools/lat_b1.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cx = bench.Ctx(); cx.dev = torch.device('cuda', 0); torch.cuda.set_device(0)
for kind in ('plain', 'full'):
    for hw in ((45, 80), (31, 56)):
        r, m = bench.latency_b1(cx, kind, hw[0], hw[1], iters=100)
        print(kind, hw, 'ms/frame median %.4f min %.4f' % (r['ms_per_frame_median'], r['ms_per_frame_min']), flush=True)
        del m
